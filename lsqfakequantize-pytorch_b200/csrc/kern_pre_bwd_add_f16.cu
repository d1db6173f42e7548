// kern_pre_bwd_add_f16.cu -- backward kernels, fused prologue M_FP32_ADD, __half tensors (see kern_pre_bwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD
#define LSQ_PRE_T __half
#define LSQ_PRE_SUFFIX add_f16
#define LSQ_PRE_MINB kMinBlocksBwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_bwd.inc"
