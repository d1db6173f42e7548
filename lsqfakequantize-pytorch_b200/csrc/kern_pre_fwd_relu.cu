// kern_pre_fwd_relu.cu -- forward kernels, fused prologue M_FP32_RELU (see kern_pre_fwd.inc).
#define LSQ_PRE_MODE M_FP32_RELU
#define LSQ_PRE_SUFFIX relu
#define LSQ_PRE_MINB kMinBlocksFwd
#define LSQ_PRE_COLUMN 1
#include "kern_pre_fwd.inc"
