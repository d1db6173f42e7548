// kern_pre_bwd_add_bf16.cu -- backward kernels, fused prologue M_FP32_ADD, __nv_bfloat16 tensors (see kern_pre_bwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD
#define LSQ_PRE_T __nv_bfloat16
#define LSQ_PRE_SUFFIX add_bf16
#define LSQ_PRE_MINB kMinBlocksBwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_bwd.inc"
