// kern_pre_bwd_addrelu_f32.cu -- backward kernels, fused prologue M_FP32_ADD_RELU, float tensors (see kern_pre_bwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD_RELU
#define LSQ_PRE_T float
#define LSQ_PRE_SUFFIX addrelu_f32
#define LSQ_PRE_MINB kMinBlocksBwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_bwd.inc"
