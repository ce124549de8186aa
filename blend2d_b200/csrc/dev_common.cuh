// dev_common.cuh - shared definitions for the device code of libb2dgpu.
//
// Every function in the dev_*.cuh headers is a small scalar routine marked B2D_HD so that the SAME source is
// compiled (a) by nvcc into the sm_100a kernels and (b) by g++ into tests/hostsim (a test-only lockstep simulator that
// lets the scalar pieces be checked against the reference on a machine without a GPU).  The product library never
// contains the host instantiation.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#  define B2D_HD __host__ __device__ __forceinline__
#  define B2D_D  __device__ __forceinline__
// Cold paths (rare operators, dithering, patterns, 64-bit dividers) are kept OUT of line on the device: the tile
// compositor's hot loop has to fit the SM's 32 KB L1.5 instruction cache (ncu round 1: 43 % of its stall samples
// were stall_no_inst while everything was inlined into an 11.7 K-instruction kernel).
#  define B2D_HD_COLD inline __host__ __device__ __noinline__
#else
#  define B2D_HD inline
#  define B2D_D  inline
#  define B2D_HD_COLD inline
#endif

namespace b2d {

// A8Info (blend2d/pipeline/pipedefs_p.h:53-59): 24.8 fixed point for coordinates, 8-bit coverage.
enum : int { kA8Shift = 8, kA8Scale = 256, kA8Mask = 255 };

// Tile geometry of the compositor (see DESIGN.md "Data layout in HBM").
enum : int {
  kTileW = 128,            // pixels per tile row  (one warp, 4 px per lane; 256 = 8 px per lane in two halves)
  kTileH = 32,             // rows per tile        (one warp per row; 32 warps = one 1024-thread CTA per SM)
  kTileThreads = 32 * kTileH
};

template<typename T> B2D_HD T tmin(T a, T b) { return b < a ? b : a; }   // bl_min (core/api.h:1742)
template<typename T> B2D_HD T tmax(T a, T b) { return a < b ? b : a; }   // bl_max (core/api.h:1747)
template<typename T> B2D_HD T tclamp(T a, T lo, T hi) { return tmin(hi, tmax(lo, a)); } // bl_clamp (:1752)

} // namespace b2d
