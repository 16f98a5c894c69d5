// marching-kernel pipeline, fp64 instantiation
#define OF3D_FAST_T double
#include "pipeline_fast.inc"
