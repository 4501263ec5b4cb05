// mi3d_internal.cuh -- declarations shared between translation units of libmi3d.so (NOT part of the C ABI)
#pragma once
#include "mi3d_common.cuh"
#include "../../include/mi3d.h"

// device-resident control block of the evaluation renderer's alive-ray loop (render.cu)
struct Mi3dEvalCtl { int n_alive, n_step, step, rows, next, done, cur, pad; };

// raymarch.cu: inference march of ctl->n_alive rays x ctl->n_step samples (extents read on the device), unused slots zeroed,
// counter[0] = ctl->rows for the field kernel that follows
int mi3d_internal_eval_march(const Mi3dEvalCtl* ctl, const int* alive, const float* rays_t, const mi3d_render_eval_args* a, const float* fars,
                             float* xyzs, float* dirs, float* deltas, int* counter, cudaStream_t st);
