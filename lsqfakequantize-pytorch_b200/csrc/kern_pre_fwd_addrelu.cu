// kern_pre_fwd_addrelu.cu -- forward kernels, fused prologue M_FP32_ADD_RELU (see kern_pre_fwd.inc).
#define LSQ_PRE_MODE M_FP32_ADD_RELU
#define LSQ_PRE_SUFFIX addrelu
#define LSQ_PRE_MINB kMinBlocksFwdAdd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_fwd.inc"
