// marching-kernel pipeline, fp32 instantiation
#define OF3D_FAST_T float
#include "pipeline_fast.inc"
