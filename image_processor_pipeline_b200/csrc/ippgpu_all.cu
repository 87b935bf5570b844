// ippgpu_all.cu — libippgpu is built as ONE translation unit (no relocatable device
// code needed for the shared device tables; whole-program inlining for the kernels).
#include "util.cu"
#include "remap.cu"
#include "rotate.cu"
#include "resample.cu"
#include "paste.cu"
#include "hsv.cu"
#include "ccl.cu"
#include "boxes.cu"
#include "chain.cu"
#include "batch.cu"
#include "enhance.cu"
#include "planner.cu"
