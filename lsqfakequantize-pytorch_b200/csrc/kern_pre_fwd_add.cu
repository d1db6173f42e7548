// kern_pre_fwd_add.cu -- forward kernels, fused prologue M_FP32_ADD (see kern_pre_fwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD
#define LSQ_PRE_SUFFIX add
#define LSQ_PRE_MINB kMinBlocksFwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_fwd.inc"
