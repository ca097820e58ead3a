// Tiled 3D step kernel for float64 (constant and variable density) on sm_100a.
//
// simwave's own 3D benchmark script builds its model in float64
// (benchmark/overthrust_3D.py:82), so the double-precision variant is what an
// unmodified run of it lands on.  Same decomposition as the float32 kernel
// (sw_step_tiled3d.cuh): a CTA owns a (BX x BY) tile of the (M,F) plane and
// marches along S over a chunk of planes; one producer warp feeds shared
// memory with TMA (u_cur planes with halo into a ring of R+1+PF slots; u_prev,
// c0 and -- only where the damping profile is non-zero -- q as halo-free tiles
// into a PS-deep ring), completion and hand-back on mbarriers, no CTA-wide
// barrier in the plane loop.  Each consumer thread updates two consecutive F
// points per plane: the F window of the pair sits in registers (128-bit shared
// loads), M neighbours are read from the centre plane in the ring, S
// neighbours live in a register queue of 2R+1 values per point.
//
// There is no two-wide float64 arithmetic to exploit and the kernel is bound
// by HBM at 40 B per point long before the FP64 pipe matters, so the update
// itself is simply value_from_neighbours<double, 3, ...> -- the one statement
// of the reference's section-1 arithmetic every kernel goes through -- on a
// neighbour accessor over the ring and the queue: bit-identical to the plain
// kernel in either math mode by construction.  Variable density: as in the
// float32 kernel the three first derivatives of the density (constant in time,
// rho_gradient_kernel) arrive as halo-free stream tiles next to u_prev, c0, q.
#pragma once

#include "sw_step_tiled3d.cuh"

namespace sw {

template <int R, int TX, int TY, int PF, int PS, bool VARDEN = false, bool RHO = VARDEN>
struct Tile3D64 {
    static constexpr int VW = 2;                     // points per thread along F
    // stream tiles of one stage: prev | c0 | q [| frF | frM | frS [| rho]] (rho
    // itself only in STRICT mode: FAST folds 1/rho into the derivatives)
    static constexpr int NSTR = VARDEN ? (RHO ? 7 : 6) : 3;
    static constexpr int RP = (R + 1) / 2 * 2;       // F halo rounded to a double2
    static constexpr int BX = TY;                    // rows (M) per tile
    static constexpr int BY = TX * VW;               // columns (F) per tile
    static constexpr int BXH = BX + 2 * R;
    static constexpr int BYH = BY + 2 * RP;
    static constexpr int NS = R + 1 + PF;            // u_cur ring slots
    static constexpr int NT = PS;                    // stream stages
    static constexpr int BOX_BYTES = BXH * BYH * 8;
    static constexpr int SLOT_BYTES = (BOX_BYTES + 127) / 128 * 128;
    static constexpr int SLOT_ELEMS = SLOT_BYTES / 8;
    static constexpr int STR_BYTES = BX * BY * 8;
    static constexpr int STR_ELEMS = BX * BY;
    static constexpr int STAGE_ELEMS = NSTR * STR_ELEMS;
    static constexpr int RING_BYTES = NS * SLOT_BYTES;
    static constexpr int STREAM_BYTES = NT * NSTR * STR_BYTES;
    static constexpr int NBARS = 2 * NS + 2 * NT;
    static constexpr int SMEM_BYTES = RING_BYTES + STREAM_BYTES + NBARS * 8 + NT * 4;
    static constexpr int CONSUMERS = TX * TY;
    static constexpr int THREADS = CONSUMERS + 32;
    static_assert(STR_BYTES % 128 == 0, "stream tiles must keep 128-byte alignment");
    static_assert(CONSUMERS % 32 == 0, "whole consumer warps");
};

// neighbourhood of one point: F window in registers, M from the centre plane
// in shared memory, S from the register queue
template <int R, int PITCH>
struct Tile64Neighbours {
    const double *w;      // &window[RP + c]: F neighbours at w[k]
    const double *ctr;    // the point in the centre plane: M neighbours at ctr[k * PITCH]
    const double *q;      // &queue[R]: S neighbours at q[k]
    __device__ __forceinline__ double C() const { return q[0]; }
    __device__ __forceinline__ double F(int k) const { return w[k]; }
    __device__ __forceinline__ double M(int k) const { return ctr[k * PITCH]; }
    __device__ __forceinline__ double M1(int k) const { return M(k); }
    __device__ __forceinline__ double S(int k) const { return q[k]; }
};

// the density of one point as the tiled kernels stream it: its value and its
// three first derivatives, summed once per run by rho_gradient_kernel
struct Tile64Density {
    static constexpr bool kDerivatives = true;
    double rho, gF, gM, gS;
    __device__ __forceinline__ double C() const { return rho; }
    __device__ __forceinline__ double frF() const { return gF; }
    __device__ __forceinline__ double frM() const { return gM; }
    __device__ __forceinline__ double frS() const { return gS; }
};

template <int R, int TX, int TY, int PF, int PS, int MATH, int MINB, bool VARDEN>
__global__ void __launch_bounds__(TX *TY + 32, MINB)
step3d_tiled64_kernel(const __grid_constant__ StepArgs<double> a,
                      const __grid_constant__ StepMaps maps,
                      const unsigned char *__restrict__ qflags, int zChunk)
{
    constexpr bool RHO = VARDEN && MATH == MATH_STRICT;
    using TL = Tile3D64<R, TX, TY, PF, PS, VARDEN, RHO>;
    constexpr int RP = TL::RP, BYH = TL::BYH, NS = TL::NS, NT = TL::NT;
    constexpr int Q = 2 * R + 1;
    constexpr int NCW = TL::CONSUMERS / 32;
    const Grid &g = a.g;

    extern __shared__ __align__(128) unsigned char smem[];
    double *ring = reinterpret_cast<double *>(smem);
    double *streams = reinterpret_cast<double *>(smem + TL::RING_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TL::RING_BYTES + TL::STREAM_BYTES);
    uint64_t *fullCur = bars, *emptyCur = bars + NS;
    uint64_t *fullStr = bars + 2 * NS, *emptyStr = bars + 2 * NS + NT;
    int *stageHasQ = reinterpret_cast<int *>(bars + TL::NBARS);

    const int tid = threadIdx.x;
    const int f0 = R + blockIdx.x * TL::BY;
    const int m0 = R + blockIdx.y * TL::BX;
    const int z0 = R + blockIdx.z * zChunk;
    const int z1 = min(z0 + zChunk, g.nS - R);
    const int planes = z1 - z0;

    if (tid == 0) {
        if (smem_u32(ring) & 127u)
            __trap();
        for (int s = 0; s < NS; s++) {
            mbar_init(&fullCur[s], 1);
            mbar_init(&emptyCur[s], NCW);
        }
        for (int s = 0; s < NT; s++) {
            mbar_init(&fullStr[s], 1);
            mbar_init(&emptyStr[s], NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // =========================== producer warp ===============================
    if (tid >= TL::CONSUMERS) {
        const int lane = tid - TL::CONSUMERS;
        const long long tilesPerPlane = (long long)gridDim.x * gridDim.y;
        const unsigned char *myFlags = qflags + (long long)blockIdx.y * gridDim.x + blockIdx.x;
        auto issue_cur = [&](int l) {
            const int slot = l % NS;
            if (l >= NS)
                mbar_wait(&emptyCur[slot], ((l / NS) - 1) & 1);
            mbar_expect_tx(&fullCur[slot], TL::BOX_BYTES);
            tma_load_3d(ring + slot * TL::SLOT_ELEMS, &maps.cur, &fullCur[slot],
                        g.lpad + f0 - RP, m0 - R, z0 - R + l);
        };
        auto issue_streams = [&](int j, int hasQ) {
            const int st = j % NT;
            if (j >= NT)
                mbar_wait(&emptyStr[st], ((j / NT) - 1) & 1);
            double *dst = streams + st * TL::STAGE_ELEMS;
            stageHasQ[st] = hasQ;
            mbar_expect_tx(&fullStr[st], ((hasQ ? 3 : 2) + (TL::NSTR - 3)) * TL::STR_BYTES);
            auto load = [&](int tile, const CUtensorMap *map) {
                tma_load_3d(dst + tile * TL::STR_ELEMS, map, &fullStr[st], g.lpad + f0, m0, z0 + j);
            };
            load(0, &maps.prev);
            load(1, &maps.c0);
            if (hasQ)
                load(2, &maps.q);
            if (VARDEN) {
                load(3, &maps.frF);
                load(4, &maps.frM);
                load(5, &maps.frS);
                if (RHO)
                    load(6, &maps.rho);
            }
        };
        const int ahead = min(max(maps.prefetch, 0), 32);
        auto prefetch_plane = [&](int j, int hasQ) {
            tma_prefetch_3d(&maps.cur, g.lpad + f0 - RP, m0 - R, z0 + R + j);
            tma_prefetch_3d(&maps.prev, g.lpad + f0, m0, z0 + j);
            tma_prefetch_3d(&maps.c0, g.lpad + f0, m0, z0 + j);
            if (hasQ)
                tma_prefetch_3d(&maps.q, g.lpad + f0, m0, z0 + j);
            if (VARDEN) {
                tma_prefetch_3d(&maps.frF, g.lpad + f0, m0, z0 + j);
                tma_prefetch_3d(&maps.frM, g.lpad + f0, m0, z0 + j);
                tma_prefetch_3d(&maps.frS, g.lpad + f0, m0, z0 + j);
                if (RHO)
                    tma_prefetch_3d(&maps.rho, g.lpad + f0, m0, z0 + j);
            }
        };
        auto flag_mask = [&](int b) -> unsigned {
            int flag = 0;
            if (b + lane < planes)
                flag = myFlags[(long long)(z0 + b + lane) * tilesPerPlane];
            return __ballot_sync(0xffffffffu, flag != 0);
        };
        unsigned long long bits = flag_mask(0);
        if (lane == 0) {
            for (int l = 0; l < 2 * R; l++)
                issue_cur(l);
            for (int j = 0; j < min(ahead, planes); j++)
                prefetch_plane(j, (int)((bits >> j) & 1ull));
        }
        for (int jb = 0; jb < planes; jb += 32) {
            bits |= (unsigned long long)flag_mask(jb + 32) << 32;
            if (lane == 0) {
                const int jend = min(jb + 32, planes);
                for (int j = jb; j < jend; j++) {
                    issue_streams(j, (int)((bits >> (j - jb)) & 1ull));
                    issue_cur(j + 2 * R);
                    const int jp = j + ahead;
                    if (ahead > 0 && jp < planes)
                        prefetch_plane(jp, (int)((bits >> (jp - jb)) & 1ull));
                }
            }
            bits >>= 32;
        }
        return;
    }

    // =========================== consumer warps ===============================
    const int lane = tid & 31;
    const int tx = tid % TX, ty = tid / TX;
    const int fMine = f0 + 2 * tx;
    const int m = m0 + ty;
    const int lastF = g.nF - R - 1, lastM = g.nM - R - 1, lastS = g.nS - R - 1;
    int nvalid = lastF - fMine + 1;
    nvalid = nvalid < 0 ? 0 : (nvalid > 2 ? 2 : nvalid);
    const bool rowValid = (m <= lastM) && nvalid > 0;
    const bool edgeTile = (m0 <= 2 * R) | (m0 + TL::BX - 1 >= lastM - R) | (f0 <= 2 * R) |
                          (f0 + TL::BY - 1 >= lastF - R);
    const int srow = ty + R;
    const int scol = 2 * tx + RP;

    auto lds2 = [&](const double *slot, int row, int col) {
        return *reinterpret_cast<const double2 *>(slot + row * BYH + col);
    };
    auto release = [&](uint64_t *bar) {
        __syncwarp();
        if (lane == 0)
            mbar_arrive(bar);
    };

    // register queue over S: qv[c][k] = plane (centre - R + k) of point c
    double qv[2][Q];
#pragma unroll
    for (int l = 0; l < 2 * R; l++) {
        mbar_wait(&fullCur[l % NS], (l / NS) & 1);
        const double2 v = lds2(ring + (l % NS) * TL::SLOT_ELEMS, srow, scol);
        qv[0][l + 1] = v.x;
        qv[1][l + 1] = v.y;
        if (l < R)
            release(&emptyCur[l % NS]);
    }

    // planes [jPlainLo, jPlainHi) lie clear of the S faces that carry a boundary
    // condition: there an interior tile stores one double2 per thread; every
    // other row goes through simple_store (boundary conditions, overhanging
    // columns, ghost copies on a neighbouring slab)
    int jPlainLo, jPlainHi;
    {
        const int sLo = (a.fuse_bc && a.bc[0] != 0) ? 2 * R + 1 : 0;
        const int sHi = (a.fuse_bc && a.bc[1] != 0) ? lastS - R : g.nS;
        jPlainLo = max(sLo - z0, 0);
        jPlainHi = min(sHi - z0, planes);
    }
    if (edgeTile)
        jPlainHi = jPlainLo;
    double *outRow = a.next + g.at(z0, m, fMine);

    int slotF = (2 * R) % NS, parF = ((2 * R) / NS) & 1;
    int slotC = R % NS;
    int st = 0, parS = 0;
#pragma unroll 2
    for (int j = 0; j < planes; j++) {
        const int s = z0 + j;

        mbar_wait(&fullCur[slotF], parF);
        {
            const double2 v = lds2(ring + slotF * TL::SLOT_ELEMS, srow, scol);
#pragma unroll
            for (int k = 0; k < Q - 1; k++) {
                qv[0][k] = qv[0][k + 1];
                qv[1][k] = qv[1][k + 1];
            }
            qv[0][Q - 1] = v.x;
            qv[1][Q - 1] = v.y;
        }

        mbar_wait(&fullStr[st], parS);
        const bool hasQ = stageHasQ[st] != 0;
        const double *sPrev = streams + st * TL::STAGE_ELEMS;
        const double *ctr = ring + slotC * TL::SLOT_ELEMS;

        // F window of my two points
        double w[2 + 2 * RP];
#pragma unroll
        for (int b = 0; b < (2 + 2 * RP) / 2; b++) {
            const double2 v = lds2(ctr, srow, scol - RP + 2 * b);
            w[2 * b] = v.x;
            w[2 * b + 1] = v.y;
        }
        const int off = (ty * TX + tx) * 2;
        const double2 pv = *reinterpret_cast<const double2 *>(sPrev + off);
        const double2 cv = *reinterpret_cast<const double2 *>(sPrev + TL::STR_ELEMS + off);
        double2 qd = make_double2(0.0, 0.0);
        if (hasQ)
            qd = *reinterpret_cast<const double2 *>(sPrev + 2 * TL::STR_ELEMS + off);
        double out[2];
        {
            const Tile64Neighbours<R, BYH> n0{w + RP, ctr + srow * BYH + scol, qv[0] + R};
            const Tile64Neighbours<R, BYH> n1{w + RP + 1, ctr + srow * BYH + scol + 1, qv[1] + R};
            if constexpr (VARDEN) {
                auto tile2 = [&](int t) {
                    return *reinterpret_cast<const double2 *>(sPrev + t * TL::STR_ELEMS + off);
                };
                const double2 gF = tile2(3), gM = tile2(4), gS = tile2(5);
                double2 rv = make_double2(1.0, 1.0);
                if (RHO)
                    rv = tile2(6);
                const Tile64Density d0{rv.x, gF.x, gM.x, gS.x}, d1{rv.y, gF.y, gM.y, gS.y};
                out[0] = value_from_neighbours<double, 3, true, R, MATH>(a, n0, d0, pv.x, cv.x, qd.x);
                out[1] = value_from_neighbours<double, 3, true, R, MATH>(a, n1, d1, pv.y, cv.y, qd.y);
            } else {
                out[0] = value_from_neighbours<double, 3, false, R, MATH>(a, n0, n0, pv.x, cv.x, qd.x);
                out[1] = value_from_neighbours<double, 3, false, R, MATH>(a, n1, n1, pv.y, cv.y, qd.y);
            }
        }

        release(&emptyCur[slotC]);
        release(&emptyStr[st]);
        if (++slotF == NS) { slotF = 0; parF ^= 1; }
        if (++slotC == NS) slotC = 0;
        if (++st == NT) { st = 0; parS ^= 1; }

        if (j >= jPlainLo && j < jPlainHi) {
            const double2 o = make_double2(out[0], out[1]);
            *reinterpret_cast<double2 *>(outRow) = o;
            if (double *alt = a.ghost_copy(s))
                *reinterpret_cast<double2 *>(alt + (outRow - a.next)) = o;
        } else if (rowValid) {
#pragma unroll
            for (int c = 0; c < 2; c++)
                if (c < nvalid)
                    simple_store<double, 3>(a, a.next, s, m, fMine + c, out[c]);
        }
        outRow += g.planeStride;
    }
}

}  // namespace sw
