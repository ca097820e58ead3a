// Shared host/device definitions of the simwave_b200 core library.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace sw {

constexpr int kMaxRadius = 10;  // space_order <= 20 (simwave model.py:47-50)
constexpr int kSplitMid = kMaxRadius / 2 + 1;   // index of pair k = 0 in StepArgs::c2odd / c1odd

// Axis naming used throughout: F = fastest (contiguous) axis, M = middle axis,
// S = slowest axis.  3D grids are (S,M,F) = (z,x,y); 2D grids are (M,F) = (z,x)
// with nS == 1.  The reference's per-axis orderings all read "F first, then M,
// then S" in this naming (sum of second derivatives 3d/wave.c:174 and
// 2d/wave.c:167, boundary passes 3d/wave.c:311-480 and 2d/wave.c:285-393),
// which is what lets one code path serve both dimensions.
enum Axis { AX_S = 0, AX_M = 1, AX_F = 2 };

// Device layout of every field: pitched rows.  Element (s,m,f) lives at
// base[(s*nM + m)*pitch + f]; `base` already includes the left padding that
// makes the first interior element of each row (f == r) 128-byte aligned.
struct Grid {
    int ndim;           // 2 or 3
    int nS, nM, nF;     // extents (nS == 1 in 2D)
    int r;              // stencil radius
    int lpad;           // elements before f == 0 in each row
    long long pitch;    // row pitch in elements
    long long planeStride;  // nM * pitch
    long long cells;    // nS * nM * pitch  (allocation without guards)

    __host__ __device__ long long at(int s, int m, int f) const {
        return ((long long)s * nM + m) * pitch + f;
    }
};

// Argument block of one time step (passed by value as a __grid_constant__).
template <typename T>
struct StepArgs {
    Grid g;
    const T *prev;      // U^{n-1}   (may alias next)
    const T *cur;       // U^{n}
    T *next;            // U^{n+1}
    const T *c0;        // dt^2 / slowness            (per point)
    const T *q;         // damp * dt / (2 * slowness) (per point; 0 outside layers)
    const T *rho;       // density or nullptr
    T c2[kMaxRadius + 1];   // second-derivative half stencil
    T c1[kMaxRadius + 1];   // first-derivative half stencil
    // "Split" F-axis order of the FAST 3D float32 kernels (sw_math.cuh,
    // split_f_sums): coefficient pairs of the odd-offset chains.  Entry
    // kSplitMid + k serves the aligned pair of wavefield values at offsets
    // (2k, 2k+1) from an even point: {coefficient of offset 2k-1 (what the
    // odd point of the pair sees in the first value), coefficient of offset
    // 2k+1 (what the even point sees in the second)}, zero beyond the radius;
    // c1odd carries the sign of the offset.
    alignas(2 * sizeof(T)) T c2odd[2 * (kMaxRadius / 2 + 1) + 1][2];
    alignas(2 * sizeof(T)) T c1odd[2 * (kMaxRadius / 2 + 1) + 1][2];
    T h2[3];            // squared spacing per axis (S,M,F)
    T inv_h2[3];        // 1/h2, correctly rounded (fast math mode only)
    T inv_h2_lo[3];     // 1/h2 - inv_h2 (fast math mode only)
    T four_h2[3];       // 4*h2 as the reference rounds it
    T inv_four_h2[3];   // 1/four_h2 (fast math mode only)
    int bc[6];          // S_before,S_after,M_before,M_after,F_before,F_after
    int quirk;          // 3D variable density with nx != ny: bug-compatible x strides
    int fuse_bc;        // 1: boundary conditions written by the step kernel itself
    int denseNx, denseNy;
    // Slab decomposition: the neighbours' u_next, shifted so that the local
    // element index of a cell in my outermost owned planes addresses its ghost
    // copy there (peer[0]: planes r..2r-1 -> up neighbour, peer[1]: planes
    // nS-2r..nS-r-1 -> down neighbour); nullptr = no such neighbour / no push.
    T *peer[2];

    // where the ghost copy of a cell of plane s lives, or nullptr
    __device__ __forceinline__ T *ghost_copy(int s) const
    {
        if (peer[0] != nullptr && s < 2 * g.r)
            return peer[0];
        if (peer[1] != nullptr && s >= g.nS - 2 * g.r)
            return peer[1];
        return nullptr;
    }
};

// Source / receiver tables on the device.
template <typename T>
struct PointTables {
    const unsigned long long *intervals;  // [count][2*ndim]
    const T *values;
    const unsigned long long *offsets;    // [count+1]
    int count;
};

class Error : public std::runtime_error {
public:
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

void set_last_error(const std::string &m);

#define SW_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) {                                             \
            throw ::sw::Error(std::string(#expr) + " failed: " +             \
                              cudaGetErrorString(_e) + " (" __FILE__ ":" +   \
                              std::to_string(__LINE__) + ")");               \
        }                                                                    \
    } while (0)

}  // namespace sw
