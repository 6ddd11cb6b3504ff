/*
 * ref_shim.h -- force-included into the reference's kernels.cu ONLY, so that the
 * untouched Pascal-era source builds for sm_100a: its legacy warp shuffles
 * (kernels.cu:16,24,79,375) no longer exist on sm_70+.
 */
#pragma once
#define __shfl_xor(v, m)  __shfl_xor_sync(0xffffffffu, (v), (m))
#define __shfl_up(v, d)   __shfl_up_sync(0xffffffffu, (v), (d))
#define __shfl_down(v, d) __shfl_down_sync(0xffffffffu, (v), (d))
