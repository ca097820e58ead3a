// Tiled 3D step kernel for sm_100a (float32; constant and variable density).
//
// Decomposition: a CTA owns a (BX x BY) tile of the (M,F) = (x,y) plane and
// marches along S (= z) over a chunk of planes.  The CTA is warp-specialised:
//
//   * one producer warp feeds shared memory with TMA (cp.async.bulk.tensor.3d,
//     one elected lane, completion on mbarriers):
//       - u_cur planes with their halo go into a ring of R+1+PF slots (the
//         R+1 planes between the centre plane and the newest plane, plus PF
//         planes in flight).  Out-of-range parts of a box are zero-filled by
//         the TMA unit, so edge tiles need no special loads;
//       - u_prev, c0 and -- only where the damping profile is non-zero inside
//         this tile and plane -- q go into a PS-deep ring of halo-free tiles.
//     Slots are handed back by the consumer warps through "empty" mbarriers,
//     so there is no CTA-wide barrier in the plane loop and the prefetch
//     depth costs no registers.
//   * TX*TY consumer threads each update a PM x 4 register tile per plane.
//     M- and F-direction neighbours are read from the centre plane in the
//     ring with 128-bit shared loads; the S-direction neighbours live in a
//     per-thread register queue of 2R+1 values per point (the classic 2.5D
//     scheme) fed from the newest plane in the ring.  u_next leaves with
//     128-bit global stores; the damping factors and the boundary conditions
//     are applied in the same kernel (store_with_boundaries semantics), with
//     a branch-free store for tiles and planes that touch no boundary.
//
// Variable density: the density enters the update only through its first
// derivatives, which do not change in time.  They are computed once per run
// (rho_gradient_kernel, same ring order and rounding as the plain kernel) and
// streamed as three more halo-free tiles next to rho itself, so the kernel
// needs neither a halo nor a register queue for the density; the first
// derivatives of u reuse the neighbours already loaded for the Laplacian.
//
// The arithmetic goes through the same helpers as the plain kernel, in the
// same order, so the two kernels agree bit for bit in either math mode (and
// with the reference in strict mode).
#pragma once

#include <cuda.h>

#include "sw_launch.h"
#include "sw_math.cuh"
#include "sw_points.cuh"
#include "sw_step_simple.cuh"

namespace sw {

// ---- PTX helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends the warp until the phase completes or the time hint
    // runs out; with a generous hint a waiting warp issues next to nothing
    // (the default hint made every wait a spin of SYNCS + BRA pairs: 10 % of
    // the issued instructions of the so-16 variable-density kernel)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// the same box fetched into L2 only (no shared memory, no barrier): lets the
// producer run far ahead of the rings without holding any on-chip space
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// ---- tile geometry -------------------------------------------------------------
template <int R, int PM, int TX, int TY, int PF, int PS, bool VARDEN = false, bool RHO = VARDEN>
struct Tile3D {
    // stream tiles of one stage: prev | c0 | q [| frF | frM | frS [| rho]]
    // (rho itself is streamed only in STRICT mode: FAST folds 1/rho into the
    // derivatives)
    static constexpr int NSTR = VARDEN ? (RHO ? 7 : 6) : 3;
    static constexpr int RP = (R + 3) / 4 * 4;       // F halo rounded to a float4
    static constexpr int BX = TY * PM;               // rows (M) per tile
    static constexpr int BY = TX * 4;                // columns (F) per tile
    static constexpr int BXH = BX + 2 * R;
    static constexpr int BYH = BY + 2 * RP;
    static constexpr int NS = R + 1 + PF;            // u_cur ring slots
    static constexpr int NT = PS;                    // stream stages
    static constexpr int BOX_BYTES = BXH * BYH * 4;
    static constexpr int SLOT_BYTES = (BOX_BYTES + 127) / 128 * 128;
    static constexpr int SLOT_FLOATS = SLOT_BYTES / 4;
    static constexpr int STR_BYTES = BX * BY * 4;    // one stream tile (multiple of 128)
    static constexpr int STR_FLOATS = BX * BY;
    static constexpr int STAGE_FLOATS = NSTR * STR_FLOATS;
    static constexpr int RING_BYTES = NS * SLOT_BYTES;
    static constexpr int STREAM_BYTES = NT * NSTR * STR_BYTES;
    static constexpr int NBARS = 2 * NS + 2 * NT;
    static constexpr int SMEM_BYTES = RING_BYTES + STREAM_BYTES + NBARS * 8 + NT * 4;
    static constexpr int CONSUMERS = TX * TY;
    static constexpr int THREADS = CONSUMERS + 32;
    static_assert(STR_BYTES % 128 == 0, "stream tiles must keep 128-byte alignment");
};

// Boundary-aware store of four consecutive F points (f .. f+3) of row (s,m).
// Same semantics as store_with_boundaries, vectorised where nothing special
// happens.  `nvalid` = how many of the four columns are interior points.
__device__ __forceinline__ void store_row4(const StepArgs<float> &a, int s, int m, int f,
                                           const float v[4], int nvalid, float *alt)
{
    const Grid &g = a.g;
    const int r = g.r;
    const int firstF = r, lastF = g.nF - r - 1;
    const int firstM = r, lastM = g.nM - r - 1;
    const int firstS = r, lastS = g.nS - r - 1;
    float *next = a.next;
    const long long p = g.at(s, m, f);
    // `alt`: ghost copy of this plane on the neighbouring slab (or nullptr)
    auto put = [&](long long idx, float x) {
        next[idx] = x;
        if (alt)
            alt[idx] = x;
    };

    const bool zMb = (a.bc[2] == 1) & (m == firstM);
    const bool zMa = (a.bc[3] == 1) & (m == lastM);
    const bool zSb = (a.bc[0] == 1) & (s == firstS);
    const bool zSa = (a.bc[1] == 1) & (s == lastS);
    const bool zRow = zMb | zMa | zSb | zSa;
    bool zFb[4], zFa[4];
    float own[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        zFb[c] = (a.bc[4] == 1) & (f + c == firstF);
        zFa[c] = (a.bc[5] == 1) & (f + c == lastF);
        own[c] = (zRow | zFb[c] | zFa[c]) ? 0.0f : v[c];
    }
    if (nvalid == 4) {
        const float4 o = make_float4(own[0], own[1], own[2], own[3]);
        *reinterpret_cast<float4 *>(next + p) = o;
        if (alt)
            *reinterpret_cast<float4 *>(alt + p) = o;
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < nvalid)
                put(p + c, own[c]);
    }

    const bool anyNeumann = (a.bc[0] == 2) | (a.bc[1] == 2) | (a.bc[2] == 2) | (a.bc[3] == 2) |
                            (a.bc[4] == 2) | (a.bc[5] == 2);
    if (!anyNeumann)
        return;

    // F pass
    if (a.bc[4] == 2 && f <= firstF + r) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int fc = f + c;
            if (c < nvalid && fc > firstF && fc <= firstF + r)
                put(p + c - 2 * (fc - firstF), v[c]);
        }
    }
    if (a.bc[5] == 2 && f + 3 >= lastF - r) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int fc = f + c;
            if (c < nvalid && fc < lastF && fc >= lastF - r)
                put(p + c + 2 * (lastF - fc), zFb[c] ? 0.0f : v[c]);
        }
    }
    // M pass sees the F pass' zeroing
    const bool mb = a.bc[2] == 2 && m > firstM && m <= firstM + r;
    const bool ma = a.bc[3] == 2 && m < lastM && m >= lastM - r;
    const bool sb = a.bc[0] == 2 && s > firstS && s <= firstS + r;
    const bool sa = a.bc[1] == 2 && s < lastS && s >= lastS - r;
    if (mb | ma | sb | sa) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c >= nvalid)
                continue;
            const bool zF = zFb[c] | zFa[c];
            const float vM = zF ? 0.0f : v[c];
            if (mb)
                put(p + c - 2 * (long long)(m - firstM) * g.pitch, vM);
            if (ma)
                put(p + c + 2 * (long long)(lastM - m) * g.pitch, zMb ? 0.0f : vM);
            // S mirrors exist only at outer faces: no ghost copy there
            const float vS = (zF | zMb | zMa) ? 0.0f : v[c];
            if (sb)
                next[p + c - 2 * (long long)(s - firstS) * g.planeStride] = vS;
            if (sa)
                next[p + c + 2 * (long long)(lastS - s) * g.planeStride] = zSb ? 0.0f : vS;
        }
    }
}

// The rare store: rows of planes near an S face, of planes with a ghost copy
// on a neighbouring slab, or of tiles on a Neumann M/F face.  Kept out of
// line so the unrolled plane loop stays small.
static __device__ __noinline__ void store_row_special(const StepArgs<float> &a, int s, int m, int f,
                                               float4 o, int nvalid)
{
    float *alt = a.ghost_copy(s);
    const float o4[4] = {o.x, o.y, o.z, o.w};
    if (a.fuse_bc) {
        store_row4(a, s, m, f, o4, nvalid, alt);
        return;
    }
    const long long p = a.g.at(s, m, f);
    if (nvalid == 4) {
        *reinterpret_cast<float4 *>(a.next + p) = o;
        if (alt)
            *reinterpret_cast<float4 *>(alt + p) = o;
    } else {
        for (int c = 0; c < nvalid; c++) {
            a.next[p + c] = o4[c];
            if (alt)
                alt[p + c] = o4[c];
        }
    }
}

// Section 2 of the reference loop for the four cells (s, m, f .. f+3) of one
// thread: every source whose window covers a cell adds its term, in index
// order, each add rounded on its own (3d/wave.c:208-295).  Rare path (only
// tiles and planes inside the bounding box of the windows get here).
static __device__ __noinline__ void add_sources_row4(const StepMaps &maps, int s, int m, int f,
                                                     const float c0v[4], const float qv[4],
                                                     float v[4])
{
    for (int src = 0; src < maps.src.count; src++) {
        long long wo = maps.step - 1;
        if (maps.waveletCount > 1)
            wo = (maps.step - 1) * maps.src.count + src;
        const float w = maps.wavelet[wo];
        if (w == 0.0f)
            continue;
        const Window<float, 3> win(maps.src, src);
        for (int c = 0; c < 4; c++)
            if (win.contains(s, m, f + c)) {
                const float kws = win.weight(s - win.lo[0], m - win.lo[1], f + c - win.lo[2]);
                v[c] = Ops<float>::add(v[c], source_term<float>(c0v[c], qv[c], kws, w));
            }
    }
}

template <int R, int PM, int TX, int TY, int PF, int PS, int MATH, int MINB, bool VARDEN,
          int UNR>
__global__ void __launch_bounds__(TX *TY + 32, MINB)
step3d_tiled_kernel(const __grid_constant__ StepArgs<float> a,
                    const __grid_constant__ StepMaps maps,
                    const unsigned char *__restrict__ qflags, int zChunk)
{
    using TL = Tile3D<R, PM, TX, TY, PF, PS, VARDEN, VARDEN && MATH == MATH_STRICT>;
    constexpr int RP = TL::RP, BYH = TL::BYH, NS = TL::NS, NT = TL::NT;
    constexpr int Q = 2 * R + 1;
    constexpr int NCW = TL::CONSUMERS / 32;
    const Grid &g = a.g;

    // TMA destinations must be 128-byte aligned; there is no static shared
    // memory in this kernel, so the dynamic segment starts the window
    extern __shared__ __align__(128) unsigned char smem[];
    float *ring = reinterpret_cast<float *>(smem);
    float *streams = reinterpret_cast<float *>(smem + TL::RING_BYTES);
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
        const unsigned char *myFlags =
            qflags + (long long)blockIdx.y * gridDim.x + blockIdx.x;
#ifdef SW_EXP_NOMEM
        // development probe (never built by default): every CTA loads the same
        // few tiles, which stay in L2 -- the SM side of the kernel alone
#define SW_PROBE_F R
#define SW_PROBE_M R
#define SW_PROBE_Z(z) (R + ((z) & 15))
#else
#define SW_PROBE_F f0
#define SW_PROBE_M m0
#define SW_PROBE_Z(z) (z)
#endif
        auto issue_cur = [&](int l) {
            const int slot = l % NS;
            if (l >= NS)
                mbar_wait(&emptyCur[slot], ((l / NS) - 1) & 1);
            mbar_expect_tx(&fullCur[slot], TL::BOX_BYTES);
            tma_load_3d(ring + slot * TL::SLOT_FLOATS, &maps.cur, &fullCur[slot],
                        g.lpad + SW_PROBE_F - RP, SW_PROBE_M - R, SW_PROBE_Z(z0 - R + l));
        };
        auto load_stream = [&](float *dst, const CUtensorMap *map, uint64_t *bar, int j) {
            tma_load_3d(dst, map, bar, g.lpad + SW_PROBE_F, SW_PROBE_M, SW_PROBE_Z(z0 + j));
        };
#undef SW_PROBE_F
#undef SW_PROBE_M
#undef SW_PROBE_Z
        auto issue_streams = [&](int j, int hasQ) {
            const int st = j % NT;
            if (j >= NT)
                mbar_wait(&emptyStr[st], ((j / NT) - 1) & 1);
            float *dst = streams + st * TL::STAGE_FLOATS;
            stageHasQ[st] = hasQ;
            constexpr int kDensityTiles = !VARDEN ? 0 : (MATH == MATH_STRICT ? 4 : 3);
            mbar_expect_tx(&fullStr[st], ((hasQ ? 3 : 2) + kDensityTiles) * TL::STR_BYTES);
            load_stream(dst, &maps.prev, &fullStr[st], j);
            load_stream(dst + TL::STR_FLOATS, &maps.c0, &fullStr[st], j);
            if (VARDEN) {
                if (MATH == MATH_STRICT)    // fast math folds 1/rho into the derivatives
                    load_stream(dst + 6 * TL::STR_FLOATS, &maps.rho, &fullStr[st], j);
                load_stream(dst + 3 * TL::STR_FLOATS, &maps.frF, &fullStr[st], j);
                load_stream(dst + 4 * TL::STR_FLOATS, &maps.frM, &fullStr[st], j);
                load_stream(dst + 5 * TL::STR_FLOATS, &maps.frS, &fullStr[st], j);
            }
            if (hasQ)
                load_stream(dst + 2 * TL::STR_FLOATS, &maps.q, &fullStr[st], j);
        };
        // L2 prefetches, `ahead` planes in front of the loads into the rings:
        // the DRAM latency is paid there, the ring loads then hit L2, so a few
        // stages of shared memory cover what is left
        const int ahead = min(max(maps.prefetch, 0), 32);
        auto prefetch_cur = [&](int l) {
            tma_prefetch_3d(&maps.cur, g.lpad + f0 - RP, m0 - R, z0 - R + l);
        };
        auto prefetch_streams = [&](int j, int hasQ) {
            tma_prefetch_3d(&maps.prev, g.lpad + f0, m0, z0 + j);
            tma_prefetch_3d(&maps.c0, g.lpad + f0, m0, z0 + j);
            if (VARDEN) {
                if (MATH == MATH_STRICT)
                    tma_prefetch_3d(&maps.rho, g.lpad + f0, m0, z0 + j);
                tma_prefetch_3d(&maps.frF, g.lpad + f0, m0, z0 + j);
                tma_prefetch_3d(&maps.frM, g.lpad + f0, m0, z0 + j);
                tma_prefetch_3d(&maps.frS, g.lpad + f0, m0, z0 + j);
            }
            if (hasQ)
                tma_prefetch_3d(&maps.q, g.lpad + f0, m0, z0 + j);
        };
        // damping flags of 32 planes starting at plane b, one per lane
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
            if (ahead > 0) {
                for (int l = 2 * R; l < min(2 * R + ahead, planes + 2 * R); l++)
                    prefetch_cur(l);
                for (int j = 0; j < min(ahead, planes); j++)
                    prefetch_streams(j, (int)((bits >> j) & 1ull));
            }
        }
        for (int jb = 0; jb < planes; jb += 32) {
            // bits: flags of planes jb .. jb+63
            bits |= (unsigned long long)flag_mask(jb + 32) << 32;
            if (lane == 0) {
                const int jend = min(jb + 32, planes);
                for (int j = jb; j < jend; j++) {
                    issue_streams(j, (int)((bits >> (j - jb)) & 1ull));
                    issue_cur(j + 2 * R);
                    if (ahead > 0) {
                        const int jp = j + ahead;
                        if (jp < planes) {
                            prefetch_streams(jp, (int)((bits >> (jp - jb)) & 1ull));
                            prefetch_cur(jp + 2 * R);
                        }
                    }
                }
            }
            bits >>= 32;
        }
        return;
    }

    // =========================== consumer warps ===============================
    const int lane = tid & 31;
    const int tx = tid % TX, ty = tid / TX;

    // my points: rows m0 + ty*PM + i, columns f0 + 4*tx .. +3
    const int fMine = f0 + 4 * tx;
    const int lastF = g.nF - R - 1, lastM = g.nM - R - 1, lastS = g.nS - R - 1;
    int nvalid = lastF - fMine + 1;
    nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
    bool rowValid[PM];
#pragma unroll
    for (int i = 0; i < PM; i++)
        rowValid[i] = (m0 + ty * PM + i <= lastM) && nvalid > 0;
    // does this tile touch (or hang over) an M/F face region where the
    // boundary conditions act?
    const bool edgeTile = (m0 <= 2 * R) | (m0 + TL::BX - 1 >= lastM - R) | (f0 <= 2 * R) |
                          (f0 + TL::BY - 1 >= lastF - R);

    const int srow = ty * PM + R;           // my first row inside a ring slot
    const int scol = 4 * tx + RP;           // my first column inside a ring slot

    auto slot_ptr = [&](int l) { return ring + (l % NS) * TL::SLOT_FLOATS; };
    auto lds4 = [&](const float *slot, int row, int col) {
        return *reinterpret_cast<const float4 *>(slot + row * BYH + col);
    };
    auto release = [&](uint64_t *bar) {
        __syncwarp();
        if (lane == 0)
            mbar_arrive(bar);
    };

    // register queue over S: qv[i][h][k] holds plane (centre - R + k) of the
    // point pair h (columns 2h, 2h+1) of row i
    float2 qv[PM][2][Q];

    // prime the queue with planes z0-R .. z0+R-1.  The first R of them are
    // never centre planes, so their slots go back as soon as every thread of
    // the warp has copied its values out.
#pragma unroll
    for (int l = 0; l < 2 * R; l++) {
        mbar_wait(&fullCur[l % NS], (l / NS) & 1);
        const float *slot = slot_ptr(l);
#pragma unroll
        for (int i = 0; i < PM; i++) {
            const float4 v = lds4(slot, srow + i, scol);
            // positions 1..2R: the loop's shift brings them to 0..2R-1
            qv[i][0][l + 1] = make_float2(v.x, v.y);
            qv[i][1][l + 1] = make_float2(v.z, v.w);
        }
        if (l < R)
            release(&emptyCur[l % NS]);
    }

    // Store tiers.  Planes [jPlainLo, jPlainHi) of this CTA lie clear of the S
    // faces that carry a boundary condition; inside that range an interior
    // tile stores one float4 per row (tier 0), an edge tile without Neumann
    // M/F faces masks Dirichlet cells and overhanging rows / columns itself
    // (tier 1); everything else goes through store_row_special (tier 2).
    // Planes with a ghost copy on a neighbouring slab repeat the tier 0 / 1
    // store into the neighbour's memory.
    int jPlainLo, jPlainHi;
    {
        const int sLo = (a.fuse_bc && a.bc[0] != 0) ? 2 * R + 1 : 0;     // first plane with s > 2R
        const int sHi = (a.fuse_bc && a.bc[1] != 0) ? lastS - R : g.nS;  // first with s >= lastS - R
        jPlainLo = max(sLo - z0, 0);
        jPlainHi = min(sHi - z0, planes);
    }
    const bool mfNeumann = a.fuse_bc && ((a.bc[2] == 2) | (a.bc[3] == 2) | (a.bc[4] == 2) |
                                         (a.bc[5] == 2));
    const int tileTier = !edgeTile ? 0 : (mfNeumann ? 2 : 1);
    if (tileTier == 2)
        jPlainHi = jPlainLo;
    // tier 1: bit c of keep[i] = column c of row i is an interior point that
    // no Dirichlet M/F face zeroes
    unsigned keep[PM];
#pragma unroll
    for (int i = 0; i < PM; i++) {
        const int m = m0 + ty * PM + i;
        unsigned k = 0;
        const bool zRow = a.fuse_bc && (((a.bc[2] == 1) & (m == R)) | ((a.bc[3] == 1) & (m == lastM)));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int f = fMine + c;
            const bool z = a.fuse_bc && (((a.bc[4] == 1) & (f == R)) | ((a.bc[5] == 1) & (f == lastF)));
            if (!zRow && !z)
                k |= 1u << c;
        }
        keep[i] = k;
    }
    float *outRow = a.next + g.at(z0, m0 + ty * PM, fMine);
    // does a source window reach into this tile? (CTA-uniform)
    const bool srcTile = maps.srcFused && m0 <= maps.srcHi[AX_M] &&
                         m0 + TL::BX - 1 >= maps.srcLo[AX_M] && f0 <= maps.srcHi[AX_F] &&
                         f0 + TL::BY - 1 >= maps.srcLo[AX_F] && z0 <= maps.srcHi[AX_S] &&
                         z1 - 1 >= maps.srcLo[AX_S];

    // the plane loop is unrolled UNR times so that the queue shift turns into
    // register renaming inside the unrolled body
    // ring positions of the newest plane (j + 2R), the centre plane (j + R) and
    // the stream stage (j), kept as running counters instead of j % NS etc.
    int slotF = (2 * R) % NS, parF = ((2 * R) / NS) & 1;
    int slotC = R % NS;
    int st = 0, parS = 0;
#pragma unroll UNR
    for (int j = 0; j < planes; j++) {
        const int s = z0 + j;

        // shift the queue and take the newest plane from the ring
        mbar_wait(&fullCur[slotF], parF);
        {
            const float *slot = ring + slotF * TL::SLOT_FLOATS;
#pragma unroll
            for (int i = 0; i < PM; i++) {
                const float4 v = lds4(slot, srow + i, scol);
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int k = 0; k < Q - 1; k++)
                        qv[i][h][k] = qv[i][h][k + 1];
                qv[i][0][Q - 1] = make_float2(v.x, v.y);
                qv[i][1][Q - 1] = make_float2(v.z, v.w);
            }
        }

        // this plane's streams
        mbar_wait(&fullStr[st], parS);
        const bool hasQ = stageHasQ[st] != 0;
        const float *sPrev = streams + st * TL::STAGE_FLOATS;
        const float *sC0 = sPrev + TL::STR_FLOATS;
        const float *sQ = sC0 + TL::STR_FLOATS;

        const float *ctr = ring + slotC * TL::SLOT_FLOATS;
        float2 out[PM][2];

#pragma unroll
        for (int i = 0; i < PM; i++) {
            // F direction: window of 4 + 2*RP values around my four points
            float w[4 + 2 * RP];
#pragma unroll
            for (int b = 0; b < (4 + 2 * RP) / 4; b++) {
                const float4 v = lds4(ctr, srow + i, scol - RP + 4 * b);
                w[4 * b + 0] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
            }
            // the window as aligned register pairs: we[j] = (w[2j], w[2j+1])
            constexpr int NW = (4 + 2 * RP) / 2;
            float2 we[NW];
#pragma unroll
            for (int jw = 0; jw < NW; jw++)
                we[jw] = make_float2(w[2 * jw], w[2 * jw + 1]);
            Stencil3x2<MATH> acc[2];
            float2 fpF[2], fpM[2], fpS[2];      // first derivatives of u (variable density)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                acc[h].begin(a, qv[i][h][R]);
                fpF[h] = fpM[h] = fpS[h] = make_float2(0.0f, 0.0f);
            }
            constexpr bool SPLIT = kSplitF<float, 3, MATH>;
#ifdef SW_EXP_NOMATH
            // development probe (never built by default): the memory side of the
            // kernel alone -- TMA rings, stream reads, stores -- without the
            // neighbour loads and the stencil arithmetic (=1), or with the
            // neighbour loads and one add per loaded pair (=2)
            constexpr bool kRings = false;
#if SW_EXP_NOMATH == 2
            if (true) {
#pragma unroll
                for (int b = 0; b < NW; b++)
                    acc[b & 1].sF = Pair::add(acc[b & 1].sF, we[b]);
#pragma unroll
                for (int ir = 1; ir <= R; ir++) {
                    const float4 up = lds4(ctr, srow + i + ir, scol);
                    const float4 dn = lds4(ctr, srow + i - ir, scol);
                    acc[0].sM = Pair::add(acc[0].sM, Pair::add(make_float2(up.x, up.y), make_float2(dn.x, dn.y)));
                    acc[1].sM = Pair::add(acc[1].sM, Pair::add(make_float2(up.z, up.w), make_float2(dn.z, dn.w)));
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        acc[h].sS = Pair::add(acc[h].sS, Pair::add(qv[i][h][R + ir], qv[i][h][R - ir]));
                }
            }
#endif
#else
            constexpr bool kRings = true;
#endif
            if constexpr (SPLIT && kRings) {
                // F axis in split order (sw_math.cuh, split_f_sums): every
                // aligned window pair feeds an even-offset chain in place and an
                // odd-offset chain with exchanged lanes; no misaligned pairs
                constexpr int K = (R + 1) / 2;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float2 E = acc[h].sF, O = make_float2(0.0f, 0.0f);
                    float2 dE = make_float2(0.0f, 0.0f), dO = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int k = -K; k <= K; k++) {
                        const float2 P = we[RP / 2 + h + k];
                        const int e = k < 0 ? -2 * k : 2 * k;
                        if (k != 0 && e <= R) {
                            E = Pair::fma(Pair::bc(a.c2[e]), P, E);
                            if (VARDEN)
                                dE = Pair::fma(Pair::bc(k < 0 ? -a.c1[e] : a.c1[e]), P, dE);
                        }
                        const int oy = 2 * k - 1 < 0 ? 1 - 2 * k : 2 * k - 1;
                        const int ox = 2 * k + 1 < 0 ? -1 - 2 * k : 2 * k + 1;
                        if (oy <= R || ox <= R) {
                            O = Pair::fma(*reinterpret_cast<const float2 *>(a.c2odd[kSplitMid + k]),
                                          P, O);
                            if (VARDEN)
                                dO = Pair::fma(
                                    *reinterpret_cast<const float2 *>(a.c1odd[kSplitMid + k]), P, dO);
                        }
                    }
                    acc[h].sF = Pair::add(E, make_float2(O.y, O.x));
                    fpF[h] = Pair::add(dE, make_float2(dO.y, dO.x));
                }
            }
            // misaligned pairs (odd F offsets) for the ring order of STRICT mode:
            // wo[j] = (w[2j+1], w[2j+2]) costs register moves
            auto wpair = [&](int k) {
                return (k & 1) ? make_float2(w[k], w[k + 1]) : we[k / 2];
            };
#pragma unroll
            for (int ir = 1; ir <= (kRings ? R : 0); ir++) {
                const float4 up = lds4(ctr, srow + i + ir, scol);
                const float4 dn = lds4(ctr, srow + i - ir, scol);
                const float2 upv[2] = {make_float2(up.x, up.y), make_float2(up.z, up.w)};
                const float2 dnv[2] = {make_float2(dn.x, dn.y), make_float2(dn.z, dn.w)};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int c = 2 * h;
                    if constexpr (!SPLIT) {
                        const float2 fp = wpair(RP + c + ir);
                        const float2 fm = wpair(RP + c - ir);
                        acc[h].ringF(a, ir, fp, fm);
                        if (VARDEN)
                            fpF[h] = ring_diff2<MATH>(fpF[h], a.c1[ir], fp, fm);
                    }
                    acc[h].ringM(a, ir, upv[h], dnv[h]);
                    acc[h].ringS(a, ir, qv[i][h][R + ir], qv[i][h][R - ir]);
                    if (VARDEN) {
                        fpM[h] = ring_diff2<MATH>(fpM[h], a.c1[ir], upv[h], dnv[h]);
                        fpS[h] = ring_diff2<MATH>(fpS[h], a.c1[ir], qv[i][h][R + ir],
                                                  qv[i][h][R - ir]);
                    }
                }
            }
            const int off = ((ty * PM + i) * TX + tx) * 4;
            const float4 pv = *reinterpret_cast<const float4 *>(sPrev + off);
            const float4 cv = *reinterpret_cast<const float4 *>(sC0 + off);
            const float2 pvv[2] = {make_float2(pv.x, pv.y), make_float2(pv.z, pv.w)};
            const float2 c0a[2] = {make_float2(cv.x, cv.y), make_float2(cv.z, cv.w)};
            float2 lap[2];
#pragma unroll
            for (int h = 0; h < 2; h++)
                lap[h] = acc[h].laplacian(a);
            if (VARDEN) {
                float4 rv = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                if (MATH == MATH_STRICT)
                    rv = *reinterpret_cast<const float4 *>(sPrev + 6 * TL::STR_FLOATS + off);
                const float4 gF = *reinterpret_cast<const float4 *>(sPrev + 3 * TL::STR_FLOATS + off);
                const float4 gM = *reinterpret_cast<const float4 *>(sPrev + 4 * TL::STR_FLOATS + off);
                const float4 gS = *reinterpret_cast<const float4 *>(sPrev + 5 * TL::STR_FLOATS + off);
                const float2 rva[2] = {make_float2(rv.x, rv.y), make_float2(rv.z, rv.w)};
                const float2 gFa[2] = {make_float2(gF.x, gF.y), make_float2(gF.z, gF.w)};
                const float2 gMa[2] = {make_float2(gM.x, gM.y), make_float2(gM.z, gM.w)};
                const float2 gSa[2] = {make_float2(gS.x, gS.y), make_float2(gS.z, gS.w)};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (MATH == MATH_STRICT) {
                        lap[h].x = density_term<float, 3>(lap[h].x, fpS[h].x, gSa[h].x, fpM[h].x,
                                                          gMa[h].x, fpF[h].x, gFa[h].x, a.four_h2,
                                                          rva[h].x);
                        lap[h].y = density_term<float, 3>(lap[h].y, fpS[h].y, gSa[h].y, fpM[h].y,
                                                          gMa[h].y, fpF[h].y, gFa[h].y, a.four_h2,
                                                          rva[h].y);
                    } else {
                        lap[h] = fast_density_term2(lap[h], fpS[h], gSa[h], fpM[h], gMa[h], fpF[h],
                                                    gFa[h]);
                    }
                }
            }
            if (hasQ) {
                // absorbing layer inside this tile and plane (warp-uniform)
                const float4 qd = *reinterpret_cast<const float4 *>(sQ + off);
                const float2 qa[2] = {make_float2(qd.x, qd.y), make_float2(qd.z, qd.w)};
#pragma unroll
                for (int h = 0; h < 2; h++)
                    out[i][h] = update_pair<MATH, true>(lap[h], qv[i][h][R], pvv[h], c0a[h], qa[h]);
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++)
                    out[i][h] = update_pair<MATH, false>(lap[h], qv[i][h][R], pvv[h], c0a[h],
                                                         make_float2(0.0f, 0.0f));
            }
            if (srcTile && s >= maps.srcLo[AX_S] && s <= maps.srcHi[AX_S]) {
                const int m = m0 + ty * PM + i;
                if (m >= maps.srcLo[AX_M] && m <= maps.srcHi[AX_M] &&
                    fMine <= maps.srcHi[AX_F] && fMine + 3 >= maps.srcLo[AX_F]) {
                    float qs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    if (hasQ) {
                        const float4 qd = *reinterpret_cast<const float4 *>(sQ + off);
                        qs[0] = qd.x; qs[1] = qd.y; qs[2] = qd.z; qs[3] = qd.w;
                    }
                    const float cs[4] = {cv.x, cv.y, cv.z, cv.w};
                    float v[4] = {out[i][0].x, out[i][0].y, out[i][1].x, out[i][1].y};
                    add_sources_row4(maps, s, m, fMine, cs, qs, v);
                    out[i][0] = make_float2(v[0], v[1]);
                    out[i][1] = make_float2(v[2], v[3]);
                }
            }
        }

        // this warp is done with the centre plane and the stream stage
        release(&emptyCur[slotC]);
        release(&emptyStr[st]);
        if (++slotF == NS) { slotF = 0; parF ^= 1; }
        if (++slotC == NS) slotC = 0;
        if (++st == NT) { st = 0; parS ^= 1; }

        if (j >= jPlainLo && j < jPlainHi) {
            // ghost copy of this plane on a neighbouring slab: same element
            // index in the neighbour's (pre-shifted) field
            float *alt = a.ghost_copy(s);
            float *altRow = alt ? alt + (outRow - a.next) : nullptr;
            if (tileTier == 0) {
#pragma unroll
                for (int i = 0; i < PM; i++) {
                    const float4 o = make_float4(out[i][0].x, out[i][0].y, out[i][1].x, out[i][1].y);
                    *reinterpret_cast<float4 *>(outRow + i * g.pitch) = o;
                    if (altRow)
                        *reinterpret_cast<float4 *>(altRow + i * g.pitch) = o;
                }
            } else {
#pragma unroll
                for (int i = 0; i < PM; i++) {
                    if (!rowValid[i])
                        continue;
                    float4 o;
                    o.x = (keep[i] & 1u) ? out[i][0].x : 0.0f;
                    o.y = (keep[i] & 2u) ? out[i][0].y : 0.0f;
                    o.z = (keep[i] & 4u) ? out[i][1].x : 0.0f;
                    o.w = (keep[i] & 8u) ? out[i][1].y : 0.0f;
                    float *dst = outRow + i * g.pitch;
                    float *dst2 = altRow ? altRow + i * g.pitch : nullptr;
                    if (nvalid == 4) {
                        *reinterpret_cast<float4 *>(dst) = o;
                        if (dst2)
                            *reinterpret_cast<float4 *>(dst2) = o;
                    } else {
                        if (nvalid > 0) dst[0] = o.x;
                        if (nvalid > 1) dst[1] = o.y;
                        if (nvalid > 2) dst[2] = o.z;
                        if (dst2) {
                            if (nvalid > 0) dst2[0] = o.x;
                            if (nvalid > 1) dst2[1] = o.y;
                            if (nvalid > 2) dst2[2] = o.z;
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PM; i++)
                if (rowValid[i])
                    store_row_special(a, s, m0 + ty * PM + i, fMine,
                                      make_float4(out[i][0].x, out[i][0].y, out[i][1].x,
                                                  out[i][1].y),
                                      nvalid);
        }
        outRow += g.planeStride;
    }
}

}  // namespace sw
