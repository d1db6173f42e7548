// kern_pre_bwd_addrelu_bf16.cu -- backward kernels, fused prologue M_FP32_ADD_RELU, __nv_bfloat16 tensors (see kern_pre_bwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD_RELU
#define LSQ_PRE_T __nv_bfloat16
#define LSQ_PRE_SUFFIX addrelu_bf16
#define LSQ_PRE_MINB kMinBlocksBwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_bwd.inc"
