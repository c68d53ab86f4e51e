// Hand-written sm_100a kernels of the SURF-cascade detection path.
//
//   k_strip_carry      per (frame,row): exact int32 row-prefix of the 8 gradient channels at every 32-column
//                      strip boundary (REDUX warp sums of packed channel pairs)
//   k_integral_walk    one warp per (frame,strip): fused gradient + channel + packed warp-shuffle row scan, then the
//                      reference's SEQUENTIAL float32 column recurrence, stored in the lattice-deinterleaved,
//                      half-split layout (sc_plan.h) with two 16 B stores per lane
//   k_cell_bounds      per frame: upper bound of every cell's box sums per projected cell edge (certifies the compact plane)
//   k_scan_stage0      64x16-window tiles: prefilter -> in-block compaction (ballot + shared prefix) -> stage 0 on
//                      the dense survivor list -> warp-aggregated push of survivors, multi/prefilter bitmasks
//   k_scan_stage       stages 1..N-1 on the compacted survivor index lists (ballot + atomic prefix between stages)
//   k_replay_rows      one thread per lattice row: exact replay of the reference's adaptive x stride from the bitmask
//   k_finalize         visited filter, detections, reference-equivalent work counters
//
// Arithmetic follows SURVEY.md Appendix A operation by operation: every float op is a single IEEE binary32
// rounding (__fadd_rn/__fsub_rn/__fmul_rn are never contracted into FMA), IEEE sqrt and divide, double sigmoid.
#ifndef SC_KERNELS_CUH
#define SC_KERNELS_CUH

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "sc_plan.h"

namespace sck {

#ifdef SC_EXP_NW2  // timing experiment only (wrong results): the filter stops after two weak classifiers
#define SC_EXP_NWEAK(n) min(n, 2)
#else
#define SC_EXP_NWEAK(n) (n)
#endif

// ---------------------------------------------------------------------------------------------------------
// Checked build (-DSC_CHECKED, surfcascade_b200/build.py --checked): every gather from an integral-image plane -- the corner
// fetches of the descriptors, the prefilter, the compact plane, the far-corner probes -- is tested against the address ranges of
// the integral buffers before it is issued; a violation is counted (sc_checked_violations) and the load replaced by zeros.
// compute-sanitizer is closed on this pool; this build is what tests/test_gpu_checked_build.py runs in its place.
// ---------------------------------------------------------------------------------------------------------
#ifdef SC_CHECKED
__device__ unsigned long long sc_chk_range[4];   // [lo0, hi0, lo1, hi1): the scan's integral buffer and the hooks' one
__device__ unsigned int sc_chk_bad;
__device__ __forceinline__ bool sc_chk(const void* p, unsigned bytes) {
    const unsigned long long a = (unsigned long long)p;
    const bool ok = (a >= sc_chk_range[0] && a + bytes <= sc_chk_range[1]) || (a >= sc_chk_range[2] && a + bytes <= sc_chk_range[3]);
    if (!ok) atomicAdd(&sc_chk_bad, 1u);
    return ok;
}
#define SC_LDG4(T, ptr) (sc_chk((ptr), 16u) ? __ldg(ptr) : T())
#else
#define SC_LDG4(T, ptr) __ldg(ptr)
#endif

// ---------------------------------------------------------------------------------------------------------
// Gradient channels (T2bFilter, DenseSURFFeatureExtractor.cpp:199-349) as four packed pairs:
// low 16 bits = negative part (even channel), high 16 bits = positive part (odd channel).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_halfwave(int d) {
    return (uint32_t)max(-d, 0) | ((uint32_t)max(d, 0) << 16);
}

struct Row3 { int a, b, c; };  // pixel values at xp, x, xn of one image row

__device__ __forceinline__ Row3 load_row3(const uint8_t* __restrict__ row, int xp, int x, int xn) {
    Row3 r;
    r.a = __ldg(row + xp); r.b = __ldg(row + x); r.c = __ldg(row + xn);
    return r;
}

__device__ __forceinline__ void channel_pairs(const Row3& prev, const Row3& cur, const Row3& next, uint32_t p[4]) {
    p[0] = pack_halfwave(cur.c - cur.a);    // dx  :224-254
    p[1] = pack_halfwave(next.b - prev.b);  // dy  :256-281
    p[2] = pack_halfwave(next.c - prev.a);  // du  :283-314
    p[3] = pack_halfwave(prev.c - next.a);  // dv  :316-347
}

// carry[frame][y][strip][8] = sum over columns left of the strip of each channel (exact, < 2^24 for W <= 65793)
__global__ void __launch_bounds__(128) k_strip_carry(const uint8_t* __restrict__ img, int W, int H, int n_strips, int nframes,
                                                      int* __restrict__ carry) {
    const int lane = threadIdx.x & 31;
    const int row_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row_id >= nframes * H) return;
    const int f = row_id / H, y = row_id - f * H;
    const uint8_t* base = img + (size_t)f * W * H;
    const uint8_t* r0 = base + (size_t)max(y - 1, 0) * W;
    const uint8_t* r1 = base + (size_t)y * W;
    const uint8_t* r2 = base + (size_t)min(y + 1, H - 1) * W;
    int run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int4* out = reinterpret_cast<int4*>(carry + ((size_t)row_id * n_strips) * 8);
    for (int s = 0; s < n_strips; s++) {
        if (lane == 0) {
            out[2 * s] = make_int4(run[0], run[1], run[2], run[3]);
            out[2 * s + 1] = make_int4(run[4], run[5], run[6], run[7]);
        }
        const int x = s * SC_STRIP + lane;
        uint32_t p[4] = {0, 0, 0, 0};
        if (x < W) {
            const int xp = max(x - 1, 0), xn = min(x + 1, W - 1);
            channel_pairs(load_row3(r0, xp, x, xn), load_row3(r1, xp, x, xn), load_row3(r2, xp, x, xn), p);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t t = __reduce_add_sync(0xffffffffu, p[k]);  // 32 * 255 < 2^16: the halves never carry
            run[2 * k] += (int)(t & 0xffffu);
            run[2 * k + 1] += (int)(t >> 16);
        }
    }
}

// One warp per (frame, strip).  S[y+1][x+1][c] = fl32(S[y][x+1][c] + float(rowprefix)), sequential in y
// (cv::integral u8->f32 as called at DenseSURFFeatureExtractor.cpp:75).
__global__ void __launch_bounds__(128) k_integral_walk(const uint8_t* __restrict__ img, int W, int H, int n_strips, int nframes,
                                                        const int* __restrict__ carry, float4* __restrict__ S, const ScLayout L) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= nframes * n_strips) return;
    const int f = wid / n_strips, s = wid - f * n_strips;
    const uint8_t* base = img + (size_t)f * W * H;
    const int x = s * SC_STRIP + lane;
    const bool valid = x < W;
    const int xc = min(x, W - 1), xp = max(xc - 1, 0), xn = min(xc + 1, W - 1);
    float4* Sf = S + (size_t)f * L.frame4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // this lane's integral column X = x + 1: in-plane column and the plane column residue are fixed
    const int X = x + 1, px = X / L.sx, rx = X - px * L.sx;
    if (valid) { float4* o = Sf + (size_t)rx * L.plane4 + SC_COL(px); o[0] = zero; o[SC_HI(L.hp)] = zero; o[SC_NOFF(L.hp)] = zero; }      // row Y = 0
    if (s == 0 && lane == 0)                                                                      // column X = 0
        for (int Y = 0; Y <= H; Y++) { float4* o = Sf + sc_layout_index(L, 0, Y); o[0] = zero; o[SC_HI(L.hp)] = zero; o[SC_NOFF(L.hp)] = zero; }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t nacc[4] = {0u, 0u, 0u, 0u};  // compact plane: (I[2k+1] << 16) + I[2k] mod 2^32 of the exact integer integral (sc_plan.h)
    const int4* cr = reinterpret_cast<const int4*>(carry + (((size_t)f * H) * n_strips + s) * 8);
    const size_t cr_step = (size_t)n_strips * 2;
    // Rows are processed in blocks of RB with the next block's inputs (pixel rows and strip carries) loaded a whole
    // block ahead: the walk is one dependent chain per warp, and with one row of lookahead every row waited on L2.
    constexpr int RB = SC_WALK_RB;
    Row3 prev = load_row3(base, xp, xc, xn), cur = prev;
    Row3 nx[RB];
    int4 clo[RB], chi[RB];
#pragma unroll
    for (int i = 0; i < RB; i++) {
        nx[i] = load_row3(base + (size_t)min(1 + i, H - 1) * W, xp, xc, xn);
        const int yq = min(i, H - 1);
        clo[i] = __ldg(cr + (size_t)yq * cr_step); chi[i] = __ldg(cr + (size_t)yq * cr_step + 1);
    }
    int py = 0, ry = 0;  // plane row / residue of Y = y + 1, advanced incrementally
    for (int y0 = 0; y0 < H; y0 += RB) {
        Row3 pn[RB];
        int4 plo[RB], phi[RB];
#pragma unroll
        for (int i = 0; i < RB; i++) {
            pn[i] = load_row3(base + (size_t)min(y0 + RB + 1 + i, H - 1) * W, xp, xc, xn);
            const int yq = min(y0 + RB + i, H - 1);
            plo[i] = __ldg(cr + (size_t)yq * cr_step); phi[i] = __ldg(cr + (size_t)yq * cr_step + 1);
        }
#pragma unroll
        for (int i = 0; i < RB; i++) {
            if (y0 + i < H) {
                uint32_t p[4];
                channel_pairs(prev, cur, nx[i], p);
                if (!valid) { p[0] = p[1] = p[2] = p[3] = 0; }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, p[k], d);
                        if (lane >= d) p[k] += t;
                    }
                }
                const int cy[8] = {clo[i].x, clo[i].y, clo[i].z, clo[i].w, chi[i].x, chi[i].y, chi[i].z, chi[i].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    acc[2 * k] = __fadd_rn(acc[2 * k], (float)(cy[2 * k] + (int)(p[k] & 0xffffu)));
                    acc[2 * k + 1] = __fadd_rn(acc[2 * k + 1], (float)(cy[2 * k + 1] + (int)(p[k] >> 16)));
                    nacc[k] += p[k] + (uint32_t)cy[2 * k] + ((uint32_t)cy[2 * k + 1] << 16);
                }
                if (++ry == L.sy) { ry = 0; py++; }
                if (valid) {
                    float4* o = Sf + (size_t)(ry * L.sx + rx) * L.plane4 + (size_t)py * L.ppitch + SC_COL(px);
                    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    o[SC_HI(L.hp)] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    reinterpret_cast<uint4*>(o)[SC_NOFF(L.hp)] = make_uint4(nacc[0], nacc[1], nacc[2], nacc[3]);
                }
                prev = cur; cur = nx[i];
            }
        }
#pragma unroll
        for (int i = 0; i < RB; i++) { nx[i] = pn[i]; clo[i] = plo[i]; chi[i] = phi[i]; }
    }
}

// Second form of the walk (default, SC_WALK_TILED): the same recurrence, but the row prefixes of a 16-row x 32-column tile
// are computed with the lanes turned by 90 degrees.  In k_integral_walk every row costs four 5-level shuffle scans
// (80 SHFL + 80 SEL + 80 adds per four rows: 60 % of its 190 instructions per row, and 20 of its 55 L1 data-pipe
// wavefronts); here lane (r, half) walks 16 columns of tile row r SERIALLY -- a running sum, no shuffles -- from a byte
// tile staged in shared memory and leaves the packed prefixes in shared memory, then the lanes become columns again for
// the reference's sequential float32 recurrence down the 16 rows.  Because the column phase reads its prefixes from shared
// memory, its lane -> column map is free: for the detection layout (sx = 4) lanes 8 rx .. 8 rx + 7 take the eight columns
// of plane residue rx, so a quarter-warp's 16-byte stores fill one 128-byte line (5 data-pipe wavefronts per STG.128
// instead of 16; the prefixes are stored in that lane order, so both shared-memory passes are conflict-free).
// Values, rounding order and layout are identical to k_integral_walk (tests: bit-exact against the oracle).
#define SC_WT 16
__global__ void __launch_bounds__(128) k_integral_walk_tiled(const uint8_t* __restrict__ img, int W, int H, int n_strips, int nframes,
                                                              const int* __restrict__ carry, float4* __restrict__ S, const ScLayout L) {
    __shared__ __align__(16) uint4 s_pre[4][SC_WT][33];      // packed inclusive row prefixes (4 channel pairs) in column-phase lane order
    __shared__ __align__(16) uint4 s_off[4][SC_WT];          // per tile row: total of the strip's columns 0..15
    __shared__ __align__(16) int4 s_car[4][SC_WT][2];        // strip carries of the tile's rows
    __shared__ __align__(16) uint8_t s_px[4][SC_WT + 2][40]; // rows y0-1 .. y0+16 (clamped), columns x0-1 .. x0+32 (clamped) at bytes 3 .. 36
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int wid = blockIdx.x * (blockDim.x >> 5) + wq;
    if (wid >= nframes * n_strips) return;  // whole warps leave: only __syncwarp below
    const int f = wid / n_strips, s = wid - f * n_strips;
    const uint8_t* base = img + (size_t)f * W * H;
    const int x0 = s * SC_STRIP;
    float4* Sf = S + (size_t)f * L.frame4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // column phase: this lane's strip column c (integral column X = x0 + c + 1)
    int c = lane;
    if (L.sx == 4) {  // u = c + 1 = 4 q + rx: lanes 8 rx + k hold q = k (+1 for rx == 0), consecutive float4s of plane column residue rx
        const int rxl = lane >> 3, k = lane & 7;
        c = 4 * (k + (rxl == 0 ? 1 : 0)) + rxl - 1;
    }
    const int x = x0 + c;
    const bool valid = x < W;
    const int X = x + 1, px = X / L.sx, rx = X - px * L.sx;
    const bool upper = c >= 16;
    if (valid) { float4* o = Sf + (size_t)rx * L.plane4 + SC_COL(px); o[0] = zero; o[SC_HI(L.hp)] = zero; o[SC_NOFF(L.hp)] = zero; }      // row Y = 0
    if (s == 0 && lane == 0)                                                                      // column X = 0
        for (int Y = 0; Y <= H; Y++) { float4* o = Sf + sc_layout_index(L, 0, Y); o[0] = zero; o[SC_HI(L.hp)] = zero; o[SC_NOFF(L.hp)] = zero; }
    // row phase: lane (r, half) owns tile row r, strip columns 16 half .. 16 half + 15
    const int r = lane & 15, half = lane >> 4;
    // position of strip column cc in the column-phase lane order (identity unless sx == 4)
    auto slot_of = [&](int cc) -> int {
        if (L.sx != 4) return cc;
        const int u = cc + 1, rxu = u & 3, q = u >> 2;
        return 8 * rxu + q - (rxu == 0 ? 1 : 0);
    };
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t nacc[4] = {0u, 0u, 0u, 0u};  // compact plane (sc_plan.h), as in k_integral_walk
    const int4* cr = reinterpret_cast<const int4*>(carry + (((size_t)f * H) * n_strips + s) * 8);
    const size_t cr_step = (size_t)n_strips * 2;
    int py = 0, ry = 0;  // plane row / residue of Y = y + 1, advanced incrementally
    const int xl = min(x0 + lane, W - 1);                                   // fill: this lane's image column
    const int xh = lane == 0 ? max(x0 - 1, 0) : min(x0 + 32, W - 1);         // halo columns (lanes 0 and 1)
    for (int y0 = 0; y0 < H; y0 += SC_WT) {
        // ---- stage the byte tile and the strip carries of these rows
#pragma unroll
        for (int t = 0; t < SC_WT + 2; t++) {
            const uint8_t* row = base + (size_t)min(max(y0 - 1 + t, 0), H - 1) * W;
            s_px[wq][t][4 + lane] = __ldg(row + xl);
            if (lane < 2) s_px[wq][t][lane == 0 ? 3 : 36] = __ldg(row + xh);
        }
        {
            const int yq = min(y0 + (lane >> 1), H - 1);
            s_car[wq][lane >> 1][lane & 1] = __ldg(cr + (size_t)yq * cr_step + (lane & 1));
        }
        __syncwarp();
        // ---- row phase: serial packed prefix of 16 columns of tile row r
        {
            const uint32_t* pm = reinterpret_cast<const uint32_t*>(&s_px[wq][r][16 * half]);      // image row y - 1
            const uint32_t* pc = reinterpret_cast<const uint32_t*>(&s_px[wq][r + 1][16 * half]);  // y
            const uint32_t* pn = reinterpret_cast<const uint32_t*>(&s_px[wq][r + 2][16 * half]);  // y + 1
            uint32_t wm[6], wc[6], wn[6];
#pragma unroll
            for (int i = 0; i < 6; i++) { wm[i] = pm[i]; wc[i] = pc[i]; wn[i] = pn[i]; }
            // byte b of the 24 loaded bytes = strip column 16 half + b - 4
            auto byte_of = [](const uint32_t* w, int b) -> int { return (int)((w[b >> 2] >> (8 * (b & 3))) & 0xffu); };
            Row3 prev, cur, next;
            prev.a = byte_of(wm, 3); prev.b = byte_of(wm, 4);
            cur.a = byte_of(wc, 3); cur.b = byte_of(wc, 4);
            next.a = byte_of(wn, 3); next.b = byte_of(wn, 4);
            uint4 run = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                prev.c = byte_of(wm, 5 + k); cur.c = byte_of(wc, 5 + k); next.c = byte_of(wn, 5 + k);
                uint32_t p[4];
                channel_pairs(prev, cur, next, p);
                run.x += p[0]; run.y += p[1]; run.z += p[2]; run.w += p[3];
                s_pre[wq][r][slot_of(16 * half + k)] = run;
                prev.a = prev.b; prev.b = prev.c; cur.a = cur.b; cur.b = cur.c; next.a = next.b; next.b = next.c;
            }
            if (half == 0) s_off[wq][r] = run;
        }
        __syncwarp();
        // ---- column phase: the reference's sequential recurrence down the tile's rows
        const int nrow = min(SC_WT, H - y0);
        for (int i = 0; i < nrow; i++) {
            uint4 pr = s_pre[wq][i][lane];
            if (upper) { const uint4 o = s_off[wq][i]; pr.x += o.x; pr.y += o.y; pr.z += o.z; pr.w += o.w; }
            const int4 clo = s_car[wq][i][0], chi = s_car[wq][i][1];
            const uint32_t p[4] = {pr.x, pr.y, pr.z, pr.w};
            const int cy[8] = {clo.x, clo.y, clo.z, clo.w, chi.x, chi.y, chi.z, chi.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc[2 * k] = __fadd_rn(acc[2 * k], (float)(cy[2 * k] + (int)(p[k] & 0xffffu)));
                acc[2 * k + 1] = __fadd_rn(acc[2 * k + 1], (float)(cy[2 * k + 1] + (int)(p[k] >> 16)));
                nacc[k] += p[k] + (uint32_t)cy[2 * k] + ((uint32_t)cy[2 * k + 1] << 16);
            }
            if (++ry == L.sy) { ry = 0; py++; }
            if (valid) {
                float4* o = Sf + (size_t)(ry * L.sx + rx) * L.plane4 + (size_t)py * L.ppitch + SC_COL(px);
                o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                o[SC_HI(L.hp)] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                reinterpret_cast<uint4*>(o)[SC_NOFF(L.hp)] = make_uint4(nacc[0], nacc[1], nacc[2], nacc[3]);
            }
        }
        __syncwarp();
    }
}

// Layout -> the reference's interleaved (H+1) x (W+1) x 8 float image (parity hook output of sc_integral).
__global__ void k_export_integral(const float4* __restrict__ S, const ScLayout L, int W, int H, float4* __restrict__ out) {
    const int n = (W + 1) * (H + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int Y = i / (W + 1), X = i - Y * (W + 1);
        const float4* p = S + sc_layout_index(L, X, Y);
        out[2 * (size_t)i] = p[0];
        out[2 * (size_t)i + 1] = p[SC_HI(L.hp)];
    }
}

// Layout -> the compact integer plane in pixel order, (H+1) x (W+1) x 4 words (parity hook of sc_integral_compact).
__global__ void k_export_compact(const float4* __restrict__ S, const ScLayout L, int W, int H, uint4* __restrict__ out) {
    const int n = (W + 1) * (H + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int Y = i / (W + 1), X = i - Y * (W + 1);
        out[i] = reinterpret_cast<const uint4*>(S + sc_layout_index(L, X, Y))[SC_NOFF(L.hp)];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Descriptor, normalisation, weak classifier
// ---------------------------------------------------------------------------------------------------------
struct Px { float v[8]; };

// Corner addressing.  ScGeom / pf offsets are BYTE offsets (uint32) from the window's low-half element, so one corner
// costs one 64-bit add (IADD3 + IADD3.X); the high half sits HP float4s further (sc_plan.h) and, with HP a template
// constant, rides in the load's immediate field.  HP == 0 selects the run-time distance `hp` (parity hooks).
template <int HP>
__device__ __forceinline__ Px load_px(const char* __restrict__ base, uint32_t off, int hp) {
    const float4* p = reinterpret_cast<const float4*>(base + off);
    if (HP < 0) {  // integral image staged in shared memory, 32 contiguous bytes per pixel (k_pool_features)
        const float4 lo = p[0], hi = p[1];
        Px r;
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w; r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
        return r;
    }
#ifdef SC_EXP_HALF  // timing experiment only (wrong results): 16 bytes per corner
    const float4 lo = __ldg(p), hi = lo;
#else
    const float4 lo = SC_LDG4(float4, p), hi = SC_LDG4(float4, p + (HP ? HP : hp));
#endif
    Px r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w; r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}

// (A + D) - (B + C), CalcFeature, DenseSURFFeatureExtractor.cpp:385-412
__device__ __forceinline__ void cell_sum(const Px& A, const Px& B, const Px& C, const Px& D, float* v) {
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = __fsub_rn(__fadd_rn(A.v[c], D.v[c]), __fadd_rn(B.v[c], C.v[c]));
}

// Sum of squares in the order of Normalize's hadd chain (:427-433 / :441-451).  Unrolling
//   s = (0,0,0,eps);  s <- hadd(s, q_g) for g = 0..7;  s <- hadd(s,s) twice
// gives  ((((((((eps + c0) + c1) + c2) + c3) + c4) + c5) + c6) + c7)  with  c_g = (q0 + q1) + (q2 + q3)  of group g.
__device__ __forceinline__ float sumsq_hadd(const float* v) {
    float s = FLT_EPSILON;
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const float q0 = __fmul_rn(v[4 * g], v[4 * g]), q1 = __fmul_rn(v[4 * g + 1], v[4 * g + 1]);
        const float q2 = __fmul_rn(v[4 * g + 2], v[4 * g + 2]), q3 = __fmul_rn(v[4 * g + 3], v[4 * g + 3]);
        s = __fadd_rn(s, __fadd_rn(__fadd_rn(q0, q1), __fadd_rn(q2, q3)));
    }
    return s;
}

// CalcFeature's 32 box sums for one projected patch; `base` = address of the window's low-half element.
template <int HP>
__device__ __forceinline__ void box_sums(const char* __restrict__ base, const ScGeom& g, int hp, float* v) {
    if (g.shape == 0) {
        // 3 x 3 corner lattice, cells row-major (GetRectsFromPatch :360-377)
        Px a0 = load_px<HP>(base, g.c[0], hp), a1 = load_px<HP>(base, g.c[1], hp), a2 = load_px<HP>(base, g.c[2], hp);
        const Px b0 = load_px<HP>(base, g.c[3], hp), b1 = load_px<HP>(base, g.c[4], hp), b2 = load_px<HP>(base, g.c[5], hp);
        cell_sum(a0, a1, b0, b1, v);
        cell_sum(a1, a2, b1, b2, v + 8);
        a0 = load_px<HP>(base, g.c[6], hp); a1 = load_px<HP>(base, g.c[7], hp); a2 = load_px<HP>(base, g.c[8], hp);
        cell_sum(b0, b1, a0, a1, v + 16);
        cell_sum(b1, b2, a1, a2, v + 24);
    } else {
        // 2 x 5 corner lattice: four cells chained along the long side (4x1 wide or 1x4 tall; B and C swap roles
        // between the two, and fl(B + C) == fl(C + B))
        Px t0 = load_px<HP>(base, g.c[0], hp), u0 = load_px<HP>(base, g.c[5], hp);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const Px t1 = load_px<HP>(base, g.c[k + 1], hp), u1 = load_px<HP>(base, g.c[k + 6], hp);
            cell_sum(t0, t1, u0, u1, v + 8 * k);
            t0 = t1; u0 = u1;
        }
    }
}

// CalcFeature + Normalize (DenseSURFFeatureExtractor.cpp:379-457), bit-exact.
template <int HP>
__device__ __forceinline__ void descriptor(const char* __restrict__ base, const ScGeom& g, int hp, float* v) {
    box_sums<HP>(base, g, hp, v);
    const float theta = 0.353553385f;  // 2 / sqrtf(32.f), DenseSURFFeatureExtractor.h:36
    const float t = __fmul_rn(__fsqrt_rn(sumsq_hadd(v)), theta);
    const float t2 = -t;
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = fmaxf(fminf(v[i], t), t2);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(sumsq_hadd(v)));
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __fmul_rn(v[i], inv);
}

// LogisticRegression::Predict, LogisticRegression.cpp:46-68.  w: 32 weights (16 B aligned), wb = double(w[32]) * bias.
__device__ __forceinline__ float weak_predict(const float* v, const float* __restrict__ w, double wb) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + i);
        s0 = __fadd_rn(__fmul_rn(wv.x, v[i]), s0);
        s1 = __fadd_rn(__fmul_rn(wv.y, v[i + 1]), s1);
        s2 = __fadd_rn(__fmul_rn(wv.z, v[i + 2]), s2);
        s3 = __fadd_rn(__fmul_rn(wv.w, v[i + 3]), s3);
    }
    const float z32 = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
    const double z = (double)z32 + wb;
    return (float)(1.0 / (1.0 + exp(-z)));
}

// One weak classifier of a stage on one window: CalcFeature + Normalize + LogisticRegression::Predict.  The projected
// geometry is read as three 16-byte vectors (shared or global memory).
template <int HP>
__device__ __forceinline__ float weak_output(const char* __restrict__ base, const ScGeom* __restrict__ geom, const float* __restrict__ w,
                                             const double* __restrict__ wb, int q, int hp) {
    ScGeom g;
    const uint4* gs = reinterpret_cast<const uint4*>(geom + q);
    const uint4 g0 = gs[0], g1 = gs[1], g2 = gs[2];
    g.c[0] = g0.x; g.c[1] = g0.y; g.c[2] = g0.z; g.c[3] = g0.w; g.c[4] = g1.x; g.c[5] = g1.y; g.c[6] = g1.z; g.c[7] = g1.w;
    g.c[8] = g2.x; g.c[9] = g2.y; g.shape = g2.z; g.pad = 0;
    float v[32];
    descriptor<HP>(base, g, hp, v);
    return weak_predict(v, w + q * SC_W_PITCH, wb[q]);
}

// GentleAdaboost::Predict2 (GentleAdaboost.cpp:247-261) of one stage on one window
template <int HP>
__device__ __forceinline__ float stage_score(const char* __restrict__ base, const ScGeom* __restrict__ geom, const float* __restrict__ w,
                                             const double* __restrict__ wb, int n_weak, int hp) {
    float acc = 0.f;
    for (int q = 0; q < n_weak; q++) acc = __fadd_rn(acc, weak_output<HP>(base, geom, w, wb, q, hp));
    return __fdiv_rn(acc, (float)n_weak);
}

// DenseSURFFeatureExtractor::sum (:351-358) and the compare at ObjDetector.cpp:188; pf = byte offsets of the corners
// (0,0) (l,0) (0,l) (l,l) from the window's low-half element
__device__ __forceinline__ float window_sum(const char* __restrict__ base, const uint32_t* pf) {
    const float4 a = SC_LDG4(float4, reinterpret_cast<const float4*>(base + pf[0])), b = SC_LDG4(float4, reinterpret_cast<const float4*>(base + pf[1]));
    const float4 c = SC_LDG4(float4, reinterpret_cast<const float4*>(base + pf[2])), d = SC_LDG4(float4, reinterpret_cast<const float4*>(base + pf[3]));
    const float s0 = __fsub_rn(__fadd_rn(a.x, d.x), __fadd_rn(b.x, c.x));
    const float s1 = __fsub_rn(__fadd_rn(a.y, d.y), __fadd_rn(b.y, c.y));
    const float s2 = __fsub_rn(__fadd_rn(a.z, d.z), __fadd_rn(b.z, c.z));
    const float s3 = __fsub_rn(__fadd_rn(a.w, d.w), __fadd_rn(b.w, c.w));
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3), 0.5f);  // /2 is exact as *0.5
}

// multi == 2 for a window rejected at stage p with stage score s (ObjDetector.cpp:201,214)
__device__ __forceinline__ bool rejected_skips(float s, int p, int n_stages) {
    return ((double)s + (double)p + 1.0) / (double)n_stages < 0.5;
}

// ---------------------------------------------------------------------------------------------------------
// Certified fast filter for stage 0.
//
// 99.7 % of the windows that reach stage 0 are rejected there, and for a rejected window the reference's outputs
// depend on the stage score only through two comparisons (score < theta, and the `multi` rule of ObjDetector.cpp:214).
// fast_weak() evaluates the same real-valued function as descriptor() + weak_predict() -- on the SAME box sums, which
// are computed with the reference's own float operations -- but with packed FFMA2 / FADD2 arithmetic, a saturating-FMA
// clip, MUFU rsqrt / ex2 / rcp and a float sigmoid: ~200 instructions instead of ~690.  Its distance to the reference
// arithmetic is bounded (below), so a window whose fast sum is further than that bound from both decision thresholds
// is decided for certain; every other window (all survivors among them) is re-evaluated with the exact path.  The
// window set, the scores of survivors and every counter are therefore identical to the exact-only kernel's.
//
// Error budget per weak classifier (u = 2^-24).  Both paths start from identical v[32].  With x = clip(v) / |clip(v)|:
//   reference: sums of squares carry <= 12u relative, sqrt / divide / products 1u each -> every x_i within 30u |x_i|
//              of the real value; the 4-lane dot adds <= 11u sum |w_i x_i|                -> |dz| <= 41u |w|_2 |x|_2
//   fast     : FFMA2 sums <= 17u, rsqrt.approx 2 ulp, saturating FMA 2^-25 absolute on g in [-0.5, 0.5] with
//              |g|_2 >= 0.5                                                              -> |dz| <= 80u |w|_2
//   sigmoid  : slope <= 1/4; ex2.approx / rcp.approx / float rounding                    -> |dp| <= 1.5e-6 absolute
// so |p_fast - p_ref| <= 0.25 * 121u * |w|_2 + 1.5e-6.  The host doubles this (ScFastParams.lim_*), and
// tests/test_gpu_parity.py measures the actual distance on ~10^6 windows (it is ~50x smaller than the budget).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// cell_sum() on channel pairs: the same two additions and one subtraction per channel, two channels per instruction
__device__ __forceinline__ void cell_sum_p(const Px& A, const Px& B, const Px& C, const Px& D, float* v) {
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        const float2 r = sub2(add2(make_float2(A.v[c], A.v[c + 1]), make_float2(D.v[c], D.v[c + 1])),
                              add2(make_float2(B.v[c], B.v[c + 1]), make_float2(C.v[c], C.v[c + 1])));
        v[c] = r.x; v[c + 1] = r.y;
    }
}

// box_sums() with packed cell sums (bit-identical values)
template <int HP>
__device__ __forceinline__ void box_sums_p(const char* __restrict__ base, const ScGeom& g, int hp, float* v) {
    if (g.shape == 0) {
        Px a0 = load_px<HP>(base, g.c[0], hp), a1 = load_px<HP>(base, g.c[1], hp), a2 = load_px<HP>(base, g.c[2], hp);
        const Px b0 = load_px<HP>(base, g.c[3], hp), b1 = load_px<HP>(base, g.c[4], hp), b2 = load_px<HP>(base, g.c[5], hp);
        cell_sum_p(a0, a1, b0, b1, v);
        cell_sum_p(a1, a2, b1, b2, v + 8);
        a0 = load_px<HP>(base, g.c[6], hp); a1 = load_px<HP>(base, g.c[7], hp); a2 = load_px<HP>(base, g.c[8], hp);
        cell_sum_p(b0, b1, a0, a1, v + 16);
        cell_sum_p(b1, b2, a1, a2, v + 24);
    } else {
        Px t0 = load_px<HP>(base, g.c[0], hp), u0 = load_px<HP>(base, g.c[5], hp);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const Px t1 = load_px<HP>(base, g.c[k + 1], hp), u1 = load_px<HP>(base, g.c[k + 6], hp);
            cell_sum_p(t0, t1, u0, u1, v + 8 * k);
            t0 = t1; u0 = u1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Box sums from the compact integer plane (sc_plan.h): 16 bytes per corner instead of 32.
// For a cell, X = N(A) + N(D) - N(B) - N(C) (mod 2^32) holds the exact box sums of a channel pair in its two 16-bit
// fields whenever both are below 65536 (certified per frame by k_cell_bounds, see there); float(field) then equals the
// reference's fl(fl(A + D) - fl(B + C)) bit for bit as long as the integrals involved are <= 2^23 (certified per tile /
// unit by compact_far_ok()).  The fields are turned into floats without the conversion pipe: PRMT builds
// 0x4B00'xxxx = 2^23 + field, one packed subtraction removes the 2^23.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_sum_c(const uint4& A, const uint4& B, const uint4& C, const uint4& D, float* v) {
    const uint32_t x[4] = {A.x + D.x - B.x - C.x, A.y + D.y - B.y - C.y, A.z + D.z - B.z - C.z, A.w + D.w - B.w - C.w};
    const float2 bias = make_float2(8388608.f, 8388608.f);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float2 r = sub2(make_float2(__uint_as_float(__byte_perm(x[k], 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(x[k], 0x4B000000u, 0x7632))), bias);
        v[2 * k] = r.x; v[2 * k + 1] = r.y;
    }
}

template <int HP>
__device__ __forceinline__ uint4 load_n(const char* __restrict__ base, uint32_t off, int hp) {
    return SC_LDG4(uint4, reinterpret_cast<const uint4*>(base + off) + SC_NOFF(HP ? HP : hp));
}

// box_sums() from the compact plane (same values as box_sums() / box_sums_p() under the two certified conditions)
template <int HP>
__device__ __forceinline__ void box_sums_c(const char* __restrict__ base, const ScGeom& g, int hp, float* v) {
    if (g.shape == 0) {
        uint4 a0 = load_n<HP>(base, g.c[0], hp), a1 = load_n<HP>(base, g.c[1], hp), a2 = load_n<HP>(base, g.c[2], hp);
        const uint4 b0 = load_n<HP>(base, g.c[3], hp), b1 = load_n<HP>(base, g.c[4], hp), b2 = load_n<HP>(base, g.c[5], hp);
        cell_sum_c(a0, a1, b0, b1, v);
        cell_sum_c(a1, a2, b1, b2, v + 8);
        a0 = load_n<HP>(base, g.c[6], hp); a1 = load_n<HP>(base, g.c[7], hp); a2 = load_n<HP>(base, g.c[8], hp);
        cell_sum_c(b0, b1, a0, a1, v + 16);
        cell_sum_c(b1, b2, a1, a2, v + 24);
    } else {
        uint4 t0 = load_n<HP>(base, g.c[0], hp), u0 = load_n<HP>(base, g.c[5], hp);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 t1 = load_n<HP>(base, g.c[k + 1], hp), u1 = load_n<HP>(base, g.c[k + 6], hp);
            cell_sum_c(t0, t1, u0, u1, v + 8 * k);
            t0 = t1; u0 = u1;
        }
    }
}

// Every integral value at or left / above layout element `far` (the far corner of a tile's or unit's last window) is
// <= 2^23: then fl(A + D), fl(B + C) and their difference are exact for every corner the tile touches (the integrals
// are monotone in x and y), i.e. the reference's float box sums ARE the integer box sums the compact plane yields.
template <int HP>
__device__ __forceinline__ bool compact_far_ok(const float4* __restrict__ far) {
    const float4 lo = SC_LDG4(float4, far), hi = SC_LDG4(float4, far + HP);
    const float m = fmaxf(fmaxf(fmaxf(lo.x, lo.y), fmaxf(lo.z, lo.w)), fmaxf(fmaxf(hi.x, hi.y), fmaxf(hi.z, hi.w)));
    return m <= 8388608.f;
}

// Certification of the compact path, once per frame: an upper bound of EVERY ce x ce cell's box sums, for each cell edge
// `ce` the stage-0 patches of the plan project to ("items", host-built).  Every ce x ce cell at any pixel position lies
// inside the box [g i, g i + ce + g) x [g j, g j + ce + g) of a pitch-g grid, and the channels are non-negative, so the
// largest such box sum (all eight channels) bounds every cell sum.  g = max(ce / 4, SC_CERT_MIN_PITCH): the looser boxes of
// small cells (whose sums are far below the limit anyway) keep the pass cheap (~40 MB of cached corner reads per 1080p frame).
// Box sums are formed in double from the float32 integrals; where an integral has passed 2^24 its accumulated
// rounding (at most half an ulp per image row) is added as slack, so the bound stays an upper bound on any frame.
// cert[frame][item] = min(bound, 0xffffffff); a weak classifier may use the compact plane iff cert < SC_CELL_LIMIT.
#define SC_CERT_MIN_PITCH 16
#define SC_CERT_MAX_ITEMS 256
#define SC_CERT_ALWAYS 0xfffffffeu   // item code: 255 ce^2 < 65536, no frame-dependent bound needed
#define SC_CERT_NEVER 0xffffffffu    // item code: compact path not available for this (scale, weak classifier)
__global__ void __launch_bounds__(256) k_cell_bounds(const float4* __restrict__ S, const ScLayout L, int W, int H, const int* __restrict__ item_ce,
                                                      int n_items, int chunks, uint32_t* __restrict__ cert) {
    const int item = blockIdx.x / chunks, chunk = blockIdx.x - item * chunks, f = blockIdx.y;
    const int ce = item_ce[item];
    const int g = max(ce / 4, SC_CERT_MIN_PITCH), b = ce + g;
    const int nbx = (W + g - 1) / g, nby = (H + g - 1) / g;
    const float4* Sf = S + (size_t)f * L.frame4;
    double best = 0.0;
    for (int i = chunk * blockDim.x + threadIdx.x; i < nbx * nby; i += chunks * blockDim.x) {
        const int by = i / nbx, bx = i - by * nbx;
        const int x0 = bx * g, y0 = by * g, x1 = min(x0 + b, W), y1 = min(y0 + b, H);
        const float4* pa = Sf + sc_layout_index(L, x0, y0);
        const float4* pb = Sf + sc_layout_index(L, x1, y0);
        const float4* pc = Sf + sc_layout_index(L, x0, y1);
        const float4* pd = Sf + sc_layout_index(L, x1, y1);
        double m = 0.0, far = 0.0;
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const float4 A = __ldg(pa + hh * L.hp), B = __ldg(pb + hh * L.hp), C = __ldg(pc + hh * L.hp), D = __ldg(pd + hh * L.hp);
            m = fmax(m, fmax(fmax(((double)A.x + D.x) - ((double)B.x + C.x), ((double)A.y + D.y) - ((double)B.y + C.y)),
                             fmax(((double)A.z + D.z) - ((double)B.z + C.z), ((double)A.w + D.w) - ((double)B.w + C.w))));
            far = fmax(far, fmax(fmax((double)D.x, (double)D.y), fmax((double)D.z, (double)D.w)));
        }
        if (far >= 16777216.0) {  // inexact integrals: each of the four corners is off by at most (rows) x ulp(far) / 2
            int e = 0;
            frexp(far, &e);                                  // far in [2^(e-1), 2^e): ulp = 2^(e-24)
            m += 2.0 * (double)(H + 1) * ldexp(1.0, e - 24);
        }
        best = fmax(best, m);
    }
    // block maximum -> one atomicMax per CTA (values are non-negative: integer order == float order after the clamp)
    __shared__ uint32_t s_best;
    if (threadIdx.x == 0) s_best = 0u;
    __syncthreads();
    const uint32_t q = best >= 4294967295.0 ? 0xffffffffu : (uint32_t)ceil(best);
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, q);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_best, wmax);
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(&cert[(size_t)f * n_items + item], s_best);
}

// Bit q set: stage-0 weak classifier q of scale si may read the compact plane in frame slot f (per-frame cell bound).
__device__ __forceinline__ uint32_t compact_mask(const ScFastParams& fp, int si, const uint32_t* __restrict__ cert, int n_items, int f) {
    uint32_t m = 0;
    for (int q = 0; q < fp.n_weak; q++) {
        const uint32_t item = fp.geom[si][q][11];
        const bool ok = item == SC_CERT_ALWAYS || (item != SC_CERT_NEVER && cert[(size_t)f * n_items + item] < SC_CELL_LIMIT);
        m |= (ok ? 1u : 0u) << q;
    }
    return m;
}


// Normalize + LogisticRegression::Predict on box sums v, approximate arithmetic (error budget above).
// With t = theta |v| the clip is  clip(v_i) = 2t g_i,  g_i = sat(v_i / (2t) + 1/2) - 1/2  (one saturating FMA and one add);
// the common factor 2t cancels in  z = w . clip(v) / |clip(v)| = (w . g) / sqrt(|g|^2 + eps / (2t)^2).
template <typename WT>
__device__ __forceinline__ float fast_tail(const float* v, const WT& w, float wb) {
    float2 s = make_float2(FLT_EPSILON, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) s = fma2(make_float2(v[2 * i], v[2 * i + 1]), make_float2(v[2 * i], v[2 * i + 1]), s);
    const float a = __fmul_rn(rsqrt_approx(__fadd_rn(s.x, s.y)), 1.41421354f);  // 1 / (2 t) = 1 / (2 * 0.353553385 * |v|)
    float2 sg = make_float2(__fmul_rn(FLT_EPSILON, __fmul_rn(a, a)), 0.f), d = make_float2(0.f, 0.f);
    const float2 mh = make_float2(-0.5f, -0.5f);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float2 g = add2(make_float2(__saturatef(__fmaf_rn(v[2 * i], a, 0.5f)), __saturatef(__fmaf_rn(v[2 * i + 1], a, 0.5f))), mh);
        sg = fma2(g, g, sg);
        d = fma2(g, make_float2(w[2 * i], w[2 * i + 1]), d);
    }
    const float z = __fmaf_rn(__fadd_rn(d.x, d.y), rsqrt_approx(__fadd_rn(sg.x, sg.y)), wb);
    return rcp_approx(__fadd_rn(1.f, ex2_approx(__fmul_rn(z, -1.44269502f))));
}

// The same on box sums read from the compact plane: those are exact non-negative integers, so the two-sided clip is
// min(v_i, t) -- 32 FMNMX instead of 32 saturating FMAs + 16 packed adds -- and nothing needs scaling by 1 / (2t):
//   z = (w . c) / sqrt(|c|^2 + eps),  c_i = min(v_i, t),  t = theta sqrt(|v|^2 + eps)   (sums stay below 2^37).
// t carries the relative error of rsqrt.approx (2 ulp) times two roundings, like `a` above; min is exact; the two FFMA2 chains
// and the final rsqrt are the ones of fast_tail(): the same distance budget covers it.
template <typename WT>
__device__ __forceinline__ float fast_tail_nonneg(const float* v, const WT& w, float wb) {
    float2 s = make_float2(FLT_EPSILON, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) s = fma2(make_float2(v[2 * i], v[2 * i + 1]), make_float2(v[2 * i], v[2 * i + 1]), s);
    const float n2 = __fadd_rn(s.x, s.y);
    const float t = __fmul_rn(__fmul_rn(n2, rsqrt_approx(n2)), 0.353553385f);  // theta |v|
    float2 sg = make_float2(FLT_EPSILON, 0.f), d = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float2 g = make_float2(fminf(v[2 * i], t), fminf(v[2 * i + 1], t));
        sg = fma2(g, g, sg);
        d = fma2(g, make_float2(w[2 * i], w[2 * i + 1]), d);
    }
    const float z = __fmaf_rn(__fadd_rn(d.x, d.y), rsqrt_approx(__fadd_rn(sg.x, sg.y)), wb);
    return rcp_approx(__fadd_rn(1.f, ex2_approx(__fmul_rn(z, -1.44269502f))));
}

// Stage-0 sum of the fast weak outputs of one window.  NW > 0: the stage's weak-classifier count as a compile-time constant --
// the loop is unrolled, so weights are constant-bank operands of the FFMA2s and the corner offsets sit at fixed constant
// addresses (no LDC / index arithmetic per weak classifier); NW == 0: run-time count (fp.n_weak).
// MODE (uniform per tile / unit, from its compact mask): 0 = some weak classifiers on the compact plane, some on the float planes
// (run-time test per weak classifier); 1 = all on the compact plane; 2 = all on the float planes.  The specialised forms carry
// no code of the other path: in the mixed form the float path's twenty 16-byte loads in flight set the register pressure of the
// whole loop (spills under the 64-register cap), measured 0.1938 vs 0.1663 ms/frame for the same all-compact work.
enum { SC_MODE_MIXED = 0, SC_MODE_COMPACT = 1, SC_MODE_FLOAT = 2 };
template <int M> struct ScModeTag { static constexpr int value = M; };
template <int HP, int MODE>
__device__ __forceinline__ float filter_weak(const ScFastParams& fp, int si, int q, const char* __restrict__ base, bool compact) {
    ScGeom g;
#pragma unroll
    for (int k = 0; k < 10; k++) g.c[k] = fp.geom[si][q][k];
    g.shape = (int)fp.geom[si][q][10]; g.pad = 0;
    float v[32];
    if (MODE == SC_MODE_COMPACT || (MODE == SC_MODE_MIXED && compact)) {
        box_sums_c<HP>(base, g, HP, v);
        return fast_tail_nonneg(v, fp.w[q], fp.wb[q]);
    }
    box_sums_p<HP>(base, g, HP, v);
    return fast_tail(v, fp.w[q], fp.wb[q]);
}
// L1 prefetch of the NEXT weak classifier's corners of the same window (no register cost: the lines arrive while this weak
// classifier's tail runs).  Timing variant, off unless -DSC_PREFETCH_NEXT.
template <int HP, int MODE>
__device__ __forceinline__ void prefetch_weak(const ScFastParams& fp, int si, int q, const char* __restrict__ base, bool compact) {
#ifdef SC_PREFETCH_NEXT
    const int n = fp.geom[si][q][10] == 0u ? 9 : 10;
    const bool c = MODE == SC_MODE_COMPACT || (MODE == SC_MODE_MIXED && compact);
#pragma unroll
    for (int k = 0; k < 10; k++) {
        if (k < n) {
            const char* a = base + fp.geom[si][q][k] + (c ? 16 * SC_NOFF(HP) : 0);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
            if (!c) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + 16 * SC_HI(HP)));
        }
    }
#endif
}
template <int HP, int NW, int MODE>
__device__ __forceinline__ float filter_sum(const ScFastParams& fp, int si, const char* __restrict__ base, uint32_t cmask) {
    float sum = 0.f;
    if (NW > 0) {
#pragma unroll
        for (int q = 0; q < SC_EXP_NWEAK(NW); q++) {
            if (q + 1 < NW) prefetch_weak<HP, MODE>(fp, si, q + 1, base, (cmask >> (q + 1)) & 1u);
            sum = __fadd_rn(sum, filter_weak<HP, MODE>(fp, si, q, base, (cmask >> q) & 1u));
        }
    } else {
#pragma unroll 1
        for (int q = 0; q < SC_EXP_NWEAK(fp.n_weak); q++) sum = __fadd_rn(sum, filter_weak<HP, MODE>(fp, si, q, base, (cmask >> q) & 1u));
    }
    return sum;
}
__device__ __forceinline__ int filter_mode(uint32_t cmask, int n_weak) {
    return cmask == 0u ? SC_MODE_FLOAT : (cmask == (1u << n_weak) - 1u ? SC_MODE_COMPACT : SC_MODE_MIXED);
}

// ---------------------------------------------------------------------------------------------------------
// Stage 0 over the lattice, one x-parity per launch (see SC_TILE_X in sc_plan.h)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread16(uint32_t x) {  // bit i (i < 16) -> bit 2i
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// phase 0: lattice columns gx = 2j, every row.  phase 1: gx = 2j + 1, only gx >= start_odd[row].
//
// FAST: the certified fast filter (above) runs on the prefilter survivors; what it cannot decide is pushed as a live
// record and goes through the exact arithmetic of k_scan_stage(stage 0) afterwards.  Keeping the exact path out of this
// kernel lets it fit 64 registers: four CTAs (32 warps) per SM instead of three, 0.278 -> 0.254 ms/frame.  fp carries
// the weights / geometry / limits in the constant bank.  !FAST evaluates stage 0 exactly in place and ignores fp.
// ALL (exact variant only): force_all mode evaluates every stage of every window; the stages 1..N-1 then run right here,
// on the window whose corners this thread just gathered, instead of as N-1 more passes of k_scan_stage over a record per
// window (C4: 242 M records re-read three times).  The records it pushes are final.
template <int HP, bool FAST, bool ALL = false, int NW = 0>
__global__ void __launch_bounds__(SC_TILE_THREADS, FAST ? SC_FAST_MIN_CTAS : (ALL ? 2 : SC_STAGE0_MIN_CTAS)) k_scan_stage0(const __grid_constant__ ScFastParams fp, const ScPlan* __restrict__ plan, const float4* __restrict__ S,
                                                                  const ScGeom* __restrict__ geom_all, const float* __restrict__ w_all,
                                                                  const double* __restrict__ wb_all, uint32_t* __restrict__ multi_bits,
                                                                  uint32_t* __restrict__ pass_bits, ScRecord* __restrict__ rec,
                                                                  uint32_t* __restrict__ rec_count, uint32_t rec_cap, int phase,
                                                                  const int* __restrict__ start_odd, const uint32_t* __restrict__ cert, int n_items) {
    constexpr int SC_TILE_Y = FAST ? SC_TILE_Y_FAST : SC_TILE_Y_EXACT;  // tile rows of this variant (sc_plan.h)
    const uint32_t block = blockIdx.x;
    // per tile row and 32-column half: raw ballot words (bit = lane); phase C spreads them to lattice-column bits
    __shared__ uint32_t s_multi[SC_TILE_Y][2];
    __shared__ uint32_t s_pass[SC_TILE_Y][2];
    __shared__ uint16_t s_list[SC_TILE_X * SC_TILE_Y];
    __shared__ uint16_t s_list2[FAST ? SC_TILE_X * SC_TILE_Y : 1];
    __shared__ int s_start[SC_TILE_Y];
    __shared__ uint32_t s_count, s_count2, s_cmask;
    __shared__ ScScale s_sc;
    extern __shared__ __align__(16) unsigned char s_dyn[];  // stage-0 weights, wb, geometry

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int f, b, si = 0;
    if (FAST) {  // uniform-datapath version of the same lookup (constant bank)
        f = block / fp.blocks_per_frame;
        b = block - f * fp.blocks_per_frame;
        while (si + 1 < fp.n_scales && fp.block_base[si + 1] <= b) si++;
    } else {
        f = block / plan->blocks_per_frame;
        b = block - f * plan->blocks_per_frame;
        while (si + 1 < plan->n_scales && plan->sc[si + 1].block_base <= b) si++;
    }
    if (tid == 0) { s_count = 0; s_count2 = 0; s_sc = plan->sc[si]; }
    __syncthreads();
    const int nx = s_sc.nx, ny = s_sc.ny;
    const int tb = b - s_sc.block_base;
    const int ty = tb / s_sc.tiles_x, tx = tb - ty * s_sc.tiles_x;
    if (phase) {
        // odd columns are needed only from the row's first non-skipping even window on
        int need = 0;
        if (tid < SC_TILE_Y) {
            const int gy = ty * SC_TILE_Y + tid;
            const int st = gy < ny ? start_odd[(size_t)f * plan->rows_per_frame + s_sc.row_base + gy] : 0x7fffffff;
            s_start[tid] = st;
            need = st < nx && st < 2 * (tx * SC_TILE_X + SC_TILE_X);
        }
        if (!__syncthreads_or(need)) return;
    }
    const int n_weak = plan->n_weak[0], total_weak = plan->total_weak;
    constexpr int ppitch = SC_ROW_ELEMS(HP);
    const float4* lo4 = S + (size_t)f * plan->lay.frame4 + (size_t)s_sc.gy0 * ppitch;  // lattice row gy0 is row 0 of this plan
    if (FAST && tid == 32) {
        // which weak classifiers may read the compact plane in this tile (sc_plan.h): the frame's cell-sum bound per weak
        // classifier, and every integral up to the far corner of the tile's last window at most 2^23
        uint32_t cm = 0;
        const int jlast = (nx - 1 - phase) >> 1;  // last valid column index of this parity (negative: none)
        if (cert != nullptr && nx - 1 - phase >= 0 && tx * SC_TILE_X <= jlast) {
            const int jm = min(tx * SC_TILE_X + SC_TILE_X - 1, jlast), gym = min(ty * SC_TILE_Y + SC_TILE_Y - 1, ny - 1);
            const char* wbase = reinterpret_cast<const char*>(lo4 + (gym * ppitch + SC_COL(jm)));
            if (compact_far_ok<HP>(reinterpret_cast<const float4*>(wbase + s_sc.pf[phase][3]))) cm = compact_mask(fp, si, cert, n_items, f);
        }
        s_cmask = cm;
    }
    const int n_ld = ALL ? total_weak : n_weak;  // weak classifiers staged in shared memory: stage 0, or every stage
    ScGeom* sg = reinterpret_cast<ScGeom*>(s_dyn);                                                        // [n_ld] 48 B each
    float* sw = reinterpret_cast<float*>(s_dyn + (size_t)n_ld * sizeof(ScGeom));                          // [n_ld][36]
    double* swb = reinterpret_cast<double*>(s_dyn + (size_t)n_ld * (sizeof(ScGeom) + SC_W_PITCH * 4));    // [n_ld]
    if (!FAST) {
        for (int i = tid; i < n_ld * SC_W_PITCH; i += SC_TILE_THREADS) sw[i] = w_all[i];
        for (int i = tid; i < n_ld; i += SC_TILE_THREADS) {
            swb[i] = wb_all[i];
            sg[i] = geom_all[((size_t)phase * plan->n_scales + si) * total_weak + i];
        }
    }

    // phase A: prefilter + compaction of the passing windows of this tile
    {
        const float thr = s_sc.thr;
        const bool use_pf = plan->use_prefilter != 0;
        constexpr int NWARPS = SC_TILE_THREADS / 32;
        static_assert(SC_TILE_Y % NWARPS == 0 && 4 * SC_TILE_Y <= SC_TILE_THREADS, "stage-0 tile shape");
#pragma unroll
        for (int i = 0; i < 2 * (SC_TILE_Y / NWARPS); i++) {
            const int row = warp + NWARPS * (i >> 1), half = i & 1;
            const int j = tx * SC_TILE_X + half * 32 + lane;
            const int gx = 2 * j + phase, gy = ty * SC_TILE_Y + row;
            bool valid = gx < nx && gy < ny;
            if (phase) valid = valid && gx >= s_start[row];
            bool pass = false;
            if (valid) pass = use_pf ? (window_sum(reinterpret_cast<const char*>(lo4 + (gy * ppitch + SC_COL(j))), s_sc.pf[phase]) > thr) : true;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            const uint32_t fail = __ballot_sync(0xffffffffu, valid && !pass);
            uint32_t base = 0;
            if (lane == 0) {
                s_pass[row][half] = m;
                // prefilter failed -> multi = 2 (ObjDetector.cpp:216-217).  FAST: the windows that pass start with their bit
                // set too -- 99.7 % of them are rejected with multi = 2 -- and the filter CLEARS the bit of the few it
                // cannot decide or that do not skip: one shared-memory atomic per ~300 windows instead of one per window
                s_multi[row][half] = FAST ? (fail | m) : fail;
                base = atomicAdd(&s_count, __popc(m));
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) s_list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((row << 6) | (half << 5) | lane);
        }
    }
    __syncthreads();

    // phase B1: certified fast filter on the dense list; what it cannot decide is compacted into s_list2
    if (FAST) {
        const uint32_t n_pass = s_count;
#ifdef SC_EXP_NOCMASK  // timing experiment only: no compact path in the loop
        const uint32_t cmask = 0;
#elif defined(SC_EXP_ALLC)  // timing experiment only (wrong results): every weak classifier reads the compact plane
        const uint32_t cmask = 0xffu;
#else
        const uint32_t cmask = s_cmask;
#endif
        auto filter_pass = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            for (uint32_t i0 = 0; i0 < n_pass; i0 += SC_TILE_THREADS) {
                const uint32_t i = i0 + tid;
                bool undecided = false;
                uint32_t code = 0;
                if (i < n_pass) {
                    code = s_list[i];
                    const int row = code >> 6, half = (code >> 5) & 1, ln = code & 31;
                    const int j = tx * SC_TILE_X + half * 32 + ln;
                    const int gy = ty * SC_TILE_Y + row;
                    const char* base = reinterpret_cast<const char*>(lo4 + (gy * ppitch + SC_COL(j)));
                    const float sum = filter_sum<HP, NW, MODE>(fp, si, base, cmask);
                    if (!(sum < fp.lim_reject && sum < fp.lim_skip)) {  // not "rejected and skips for certain": rare
                        atomicAnd(&s_multi[row][half], ~(1u << ln));
                        undecided = !(sum < fp.lim_reject && sum >= fp.lim_noskip);
                        if (undecided && sum >= fp.lim_pass) code |= 0x8000u;  // passes stage 0 for certain: no exact stage 0 needed
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, undecided);
                if (m) {
                    uint32_t base2 = 0;
                    if (lane == 0) base2 = atomicAdd(&s_count2, __popc(m));
                    base2 = __shfl_sync(0xffffffffu, base2, 0);
                    if (undecided) s_list2[base2 + __popc(m & ((1u << lane) - 1u))] = (uint16_t)code;
                }
            }
        };
        const int mode = filter_mode(cmask, fp.n_weak);  // uniform over the tile
        if (mode == SC_MODE_COMPACT) filter_pass(ScModeTag<SC_MODE_COMPACT>());
        else if (mode == SC_MODE_FLOAT) filter_pass(ScModeTag<SC_MODE_FLOAT>());
        else filter_pass(ScModeTag<SC_MODE_MIXED>());
        __syncthreads();
    }

    // phase B: stage 0 with the reference's arithmetic on the dense list (FAST: on what the filter left undecided)
    const uint32_t count = FAST ? s_count2 : s_count;
    const uint16_t* list = FAST ? s_list2 : s_list;
    const float theta0 = plan->theta[0];
    const int n_stages = plan->n_stages;
    const bool force = plan->force_all != 0;
    for (uint32_t i0 = 0; i0 < count; i0 += SC_TILE_THREADS) {
        const uint32_t i = i0 + tid;
        bool push = false;
        ScRecord r;
        r.fs = 0; r.yx = 0; r.rej = 0; r.score = 0;
        if (i < count) {
            const uint32_t code = list[i];
            const int row = (code >> 6) & 0x1ff, half = (code >> 5) & 1, ln = code & 31;
            const int j = tx * SC_TILE_X + half * 32 + ln;
            const int gx = 2 * j + phase, gy = ty * SC_TILE_Y + row;
            float score = 0.f;
            bool rejected = false;
            int rej = (FAST && (code & 0x8000u)) ? -2 : -1;  // FAST: still alive; -1: k_scan_stage(0) decides, -2: stage 0 certainly passed
            if (!FAST) {
                const char* base = reinterpret_cast<const char*>(lo4 + (gy * ppitch + SC_COL(j)));
                score = stage_score<HP>(base, sg, sw, swb, n_weak, HP);
                rejected = score < theta0;
                if (rejected && rejected_skips(score, 0, n_stages)) atomicOr(&s_multi[row][half], 1u << ln);
                rej = rejected ? 0 : (n_stages == 1 ? 1 : -1);
                if (ALL) {
                    // every later stage is evaluated (force_all); the first rejection keeps its stage and score,
                    // exactly as k_scan_stage would leave the record after its N-1 passes
                    for (int st = 1; st < n_stages; st++) {
                        const int wbs = plan->weak_base[st];
                        const float sc = stage_score<HP>(base, sg + wbs, sw + (size_t)wbs * SC_W_PITCH, swb + wbs, plan->n_weak[st], HP);
                        if (rej < 0) {
                            score = sc;
                            if (sc < plan->theta[st]) {
                                rej = st;
                                if (rejected_skips(sc, st, n_stages)) atomicOr(&s_multi[row][half], 1u << ln);
                            } else if (st == n_stages - 1) {
                                rej = n_stages;
                            }
                        }
                    }
                }
            }
            push = !rejected || force;
            r.fs = ((uint32_t)f << 8) | (uint32_t)si;
            r.yx = ((uint32_t)gy << 16) | (uint32_t)gx;
            r.rej = rej;
            r.score = __float_as_uint(score);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, push);
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(rec_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (push) {
                const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < rec_cap) rec[slot] = r;
            }
        }
    }
    __syncthreads();

    // phase C: publish the four bitmask words of every tile row (phase 0 defines the words, phase 1 ORs its bits in)
    if (tid < SC_TILE_Y * 4) {
        const int row = tid >> 2, w = tid & 3;
        const int gy = ty * SC_TILE_Y + row, wx = tx * 4 + w;
        if (gy < ny && wx < s_sc.wpr) {
            const size_t wi = (size_t)f * plan->words_per_frame + s_sc.word_base + (size_t)gy * s_sc.wpr + wx;
            // word w of the row: lanes 16 (w & 1) .. + 15 of half w >> 1, one bit per lattice column 2 j + phase
            const uint32_t mw = spread16(s_multi[row][w >> 1] >> (16 * (w & 1))) << phase;
            const uint32_t pw = spread16(s_pass[row][w >> 1] >> (16 * (w & 1))) << phase;
            if (phase == 0) {
                multi_bits[wi] = mw;
                pass_bits[wi] = pw;
            } else {
                if (mw) atomicOr(&multi_bits[wi], mw);
                if (pw) atomicOr(&pass_bits[wi], pw);
            }
        }
    }
}

// After phase 0: per lattice row, the first even column whose window does not certainly skip (multi bit clear: it
// passed stage 0, or was rejected with (s + 1) / N >= 0.5).  The reference's chain x += multi * step stays on even
// columns up to there; odd columns can only be visited from the next one on.
// With row_chunks non-null it also counts, per row, the 32-window runs of reachable odd columns (start + 64 k + 2 lane):
// k_chunk_scan / k_chunk_fill turn the counts into the work list of k_scan_odd in (frame, scale, row) order, so that the
// warps pulling consecutive runs stay inside one frame's integral image (L2-resident).  An atomically appended list
// interleaved frames and scales and made k_scan_odd read 1.9 GB from HBM per 8 frames (ncu), 3.5 x their integral images.
__global__ void __launch_bounds__(128) k_row_events(const ScPlan* __restrict__ plan, int nframes, const uint32_t* __restrict__ multi_bits,
                                                     int* __restrict__ start_odd, unsigned long long* __restrict__ counters,
                                                     uint32_t* __restrict__ row_chunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = plan->rows_per_frame;
    if (t >= nframes * rows) return;
    const int f = t / rows, r = t - f * rows;
    int si = 0;
    while (si + 1 < plan->n_scales && plan->sc[si + 1].row_base <= r) si++;
    const int wpr = plan->sc[si].wpr, nx = plan->sc[si].nx;
    const size_t w0 = (size_t)f * plan->words_per_frame + plan->sc[si].word_base + (size_t)(r - plan->sc[si].row_base) * wpr;
    int start = 0x7fffffff;
    for (int wi = 0; wi < wpr; wi++) {
        const int nb = min(32, nx - 32 * wi);
        const uint32_t valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        const uint32_t z = ~multi_bits[w0 + wi] & 0x55555555u & valid;
        if (z) { start = 32 * wi + __ffs(z); break; }  // (ffs - 1) is the even column, + 1 the first odd one
    }
    start_odd[t] = start;
    if (start < nx) {
        const int n_odd = (nx - start + 1) / 2;
        atomicAdd(&counters[(size_t)f * SC_CNT_STRIDE + SC_CNT_EVALODD], (unsigned long long)n_odd);
        if (row_chunks) row_chunks[t] = (uint32_t)((n_odd + SC_ODD_UNIT - 1) / SC_ODD_UNIT);
    } else if (row_chunks) {
        row_chunks[t] = 0;
    }
}

// sum of every 1024-row block of row_chunks
__global__ void __launch_bounds__(1024) k_chunk_blocksum(const uint32_t* __restrict__ row_chunks, int n, uint32_t* __restrict__ blk) {
    __shared__ uint32_t wsum[32];
    const int t = blockIdx.x * 1024 + threadIdx.x;
    uint32_t v = t < n ? row_chunks[t] : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t s = __reduce_add_sync(0xffffffffu, wsum[threadIdx.x]);
        if (threadIdx.x == 0) blk[blockIdx.x] = s;
    }
}

// per 1024-row block: offset of the block (sum of the earlier blocks) + exclusive scan of its rows' counts, then every
// row writes its entries (row << 10 | k); the last block publishes the total
__global__ void __launch_bounds__(1024) k_chunk_fill(const uint32_t* __restrict__ row_chunks, int n, const uint32_t* __restrict__ blk,
                                                      uint32_t* __restrict__ chunk_list, uint32_t* __restrict__ chunk_count) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        uint32_t acc = 0;
        for (int b = lane; b < (int)blockIdx.x; b += 32) acc += blk[b];
        acc = __reduce_add_sync(0xffffffffu, acc);
        if (lane == 0) s_base = acc;
    }
    const int t = blockIdx.x * 1024 + tid;
    const uint32_t c = t < n ? row_chunks[t] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        wsum[lane] = wi - w;  // exclusive prefix of the warp sums
    }
    __syncthreads();
    const uint32_t off = s_base + wsum[warp] + incl - c;
    for (uint32_t k = 0; k < c; k++) chunk_list[off + k] = ((uint32_t)t << 10) | k;
    if (blockIdx.x == gridDim.x - 1 && tid == 1023) *chunk_count = off + c;
}

// Stage 0 on the reachable odd columns (the reference's stride reaches them only behind a row's first non-skipping
// window: ~16 % of the odd columns at 1080p, as ragged row suffixes).  Persistent warps pull units of up to SC_ODD_UNIT
// consecutive odd windows of one row from the list k_row_events built, so the work is proportional to the windows, not
// to the tiles they are scattered over.  Same arithmetic and decisions as k_scan_stage0<FAST = true>: prefilter (four
// 32-window passes, bits assembled from the ballots), survivors compacted into a per-warp list, certified fast filter on
// dense 32-window batches of that list, live records for what it leaves undecided.  The unit's bitmask words are
// collected in shared memory and ORed into the global masks once (<= 9 + 9 atomics per unit instead of one per window).
// The claim of the next unit (cursor atomic -> list entry -> row start: three dependent round trips to L2) is issued by
// lane 0 before the current unit is processed and only broadcast afterwards.
// (32-window units without compaction: 0.0495 ms/frame on C2, lanes 70 % busy in the filter.)
template <int HP, int NW = 0>
__global__ void __launch_bounds__(256, SC_STAGE0_MIN_CTAS) k_scan_odd(const __grid_constant__ ScFastParams fp, const ScPlan* __restrict__ plan, const float4* __restrict__ S,
                                                                      const ScGeom* __restrict__ geom_all, const float* __restrict__ w_all,
                                                                      const double* __restrict__ wb_all, uint32_t* __restrict__ multi_bits,
                                                                      uint32_t* __restrict__ pass_bits, ScRecord* __restrict__ rec,
                                                                      uint32_t* __restrict__ rec_count, uint32_t rec_cap, const int* __restrict__ start_odd,
                                                                      const uint32_t* __restrict__ chunk_list, const uint32_t* __restrict__ chunk_count,
                                                                      uint32_t* __restrict__ cursor, const uint32_t* __restrict__ cert, int n_items) {
    constexpr int NWARP = 8, WORDS = 2 * SC_ODD_UNIT / 32 + 2;  // a unit spans 2 * SC_ODD_UNIT lattice columns at an odd offset
    __shared__ uint8_t s_q[NWARP][SC_ODD_UNIT];
    __shared__ uint32_t s_mb[NWARP][WORDS], s_pb[NWARP][WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_chunks = *chunk_count;
    const int rows = plan->rows_per_frame;
    const bool use_pf = plan->use_prefilter != 0;
    constexpr int ppitch = SC_ROW_ELEMS(HP);
    const uint32_t lt = (1u << lane) - 1u;
    // claim of a unit, lane 0 only: index, list entry, first reachable odd column of its row
    uint32_t c = 0, e = 0;
    int st0 = 0;
    if (lane == 0) {
        c = atomicAdd(cursor, 1u);
        if (c < n_chunks) { e = chunk_list[c]; st0 = start_odd[e >> 10]; }
    }
    c = __shfl_sync(0xffffffffu, c, 0); e = __shfl_sync(0xffffffffu, e, 0); st0 = __shfl_sync(0xffffffffu, st0, 0);
    while (c < n_chunks) {
        uint32_t c2 = 0, e2 = 0;
        int st2 = 0;
        if (lane == 0) {  // next unit: in flight while this one is processed
            c2 = atomicAdd(cursor, 1u);
            if (c2 < n_chunks) { e2 = chunk_list[c2]; st2 = start_odd[e2 >> 10]; }
        }
        const int t = (int)(e >> 10), k = (int)(e & 1023u);
        const int f = t / rows, r = t - f * rows;
        int si = 0;
        while (si + 1 < fp.n_scales && fp.row_base[si + 1] <= r) si++;
        const ScScale* sc = &plan->sc[si];
        const int gy = r - fp.row_base[si], nx = sc->nx;
        const int g0 = st0 + 2 * SC_ODD_UNIT * k;  // lattice column (odd) of the unit's first window
        const float4* row4 = S + (size_t)f * plan->lay.frame4 + (size_t)(gy + sc->gy0) * ppitch;
        const size_t word0 = (size_t)f * plan->words_per_frame + sc->word_base + (size_t)gy * sc->wpr + (g0 >> 5);
        if (lane < WORDS) { s_mb[warp][lane] = 0; s_pb[warp][lane] = 0; }
        // compact-plane eligibility of this unit (as in k_scan_stage0): per-frame cell bounds, far corner of its last window
        uint32_t cmask = 0;
        if (lane == 0 && cert != nullptr) {
            const int glast = min(g0 + 2 * (SC_ODD_UNIT - 1), ((nx - 1) & 1) ? nx - 1 : nx - 2);  // last odd column of the unit
            if (glast >= g0 && glast < nx && compact_far_ok<HP>(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(row4 + SC_COL(glast >> 1)) + sc->pf[1][3])))
                cmask = compact_mask(fp, si, cert, n_items, f);
        }
        cmask = __shfl_sync(0xffffffffu, cmask, 0);
#ifdef SC_EXP_ALLC
        cmask = 0xffu;
#endif
        __syncwarp();
        // prefilter; pass / prefilter-failed bits of lane i sit at bit (gc & 31) + 2 i of the 96-bit run starting at word wb
        int n = 0;
        {
            const float thr = sc->thr;
            uint32_t pf[4];
#pragma unroll
            for (int i = 0; i < 4; i++) pf[i] = sc->pf[1][i];
#pragma unroll
            for (int ch = 0; ch < SC_ODD_UNIT / 32; ch++) {
                const int gc = g0 + 64 * ch, gx = gc + 2 * lane;
                const bool valid = gx < nx;
                bool pass = false;
                if (valid) pass = use_pf ? (window_sum(reinterpret_cast<const char*>(row4 + SC_COL(gx >> 1)), pf) > thr) : true;
                const uint32_t m = __ballot_sync(0xffffffffu, pass);
                const uint32_t fail = __ballot_sync(0xffffffffu, valid && !pass);
                if (pass) s_q[warp][n + __popc(m & lt)] = (uint8_t)(32 * ch + lane);
                n += __popc(m);
                if (lane < 2) {  // lane 0: pass bits; lane 1: multi bits = prefilter failed (ObjDetector.cpp:216-217) or passed --
                                 // the filter below clears the bit of the few survivors that are not "rejected and skip"
                    const uint32_t b = lane ? (fail | m) : m;
                    const unsigned long long v = ((unsigned long long)spread16(b) | ((unsigned long long)spread16(b >> 16) << 32));
                    const int sh = gc & 31, wb = (gc >> 5) - (g0 >> 5);  // sh is odd: never 0
                    uint32_t* dst = lane ? s_mb[warp] : s_pb[warp];
                    dst[wb] |= (uint32_t)(v << sh);
                    dst[wb + 1] |= (uint32_t)(v >> (32 - sh));
                    dst[wb + 2] |= (uint32_t)(v >> (64 - sh));
                }
            }
        }
        __syncwarp();
        // certified fast filter on dense batches of the survivors
        auto filter_pass = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            for (int b0 = 0; b0 < n; b0 += 32) {
                const int ei = b0 + lane;
                bool exact = false, passed = false;
                int gx = 0;
                if (ei < n) {
                    gx = g0 + 2 * (int)s_q[warp][ei];
                    const char* base = reinterpret_cast<const char*>(row4 + SC_COL(gx >> 1));
                    const float sum = filter_sum<HP, NW, MODE>(fp, si, base, cmask);
                    if (!(sum < fp.lim_reject && sum < fp.lim_skip)) {
                        atomicAnd(&s_mb[warp][(gx >> 5) - (g0 >> 5)], ~(1u << (gx & 31)));
                        exact = !(sum < fp.lim_reject && sum >= fp.lim_noskip);
                        passed = exact && sum >= fp.lim_pass;
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, exact);
                if (m) {  // rare: left to the exact arithmetic of k_scan_stage(stage 0) as live records
                    uint32_t slot0 = 0;
                    if (lane == 0) slot0 = atomicAdd(rec_count, __popc(m));
                    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                    if (exact) {
                        const uint32_t slot = slot0 + __popc(m & lt);
                        if (slot < rec_cap) {
                            ScRecord rc;
                            rc.fs = ((uint32_t)f << 8) | (uint32_t)si;
                            rc.yx = ((uint32_t)gy << 16) | (uint32_t)gx;
                            rc.rej = passed ? -2 : -1; rc.score = 0;  // -2: stage 0 certainly passed (no exact stage 0 needed)
                            rec[slot] = rc;
                        }
                    }
                }
            }
        };
        {
            const int mode = filter_mode(cmask, fp.n_weak);  // uniform over the unit
            if (mode == SC_MODE_COMPACT) filter_pass(ScModeTag<SC_MODE_COMPACT>());
            else if (mode == SC_MODE_FLOAT) filter_pass(ScModeTag<SC_MODE_FLOAT>());
            else filter_pass(ScModeTag<SC_MODE_MIXED>());
        }
        __syncwarp();
        if (lane < WORDS) {
            const uint32_t mb = s_mb[warp][lane], pb = s_pb[warp][lane];
            if (mb) atomicOr(&multi_bits[word0 + lane], mb);
            if (pb) atomicOr(&pass_bits[word0 + lane], pb);
        }
        __syncwarp();
        c = __shfl_sync(0xffffffffu, c2, 0); e = __shfl_sync(0xffffffffu, e2, 0); st0 = __shfl_sync(0xffffffffu, st2, 0);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stages 1..N-1 (and, behind the fast filter, the exact stage 0) on compacted index lists, one thread per window.
// (Spreading a window's weak classifiers over 2..8 lanes and adding their outputs in the reference's order through
// shuffles was measured slower, 0.0247 -> 0.0269 ms/frame on C2 and +9 % on C4: ncu shows the kernel waiting on scattered
// 32-byte sector reads -- DRAM 50 % busy at 0.35 GB per launch, L2 hit rate 35 % because the 8 integral images of a scan
// group have left L2 by then -- not on the serial chain, and fewer windows per warp lose the lines x-adjacent survivors share.)
// ---------------------------------------------------------------------------------------------------------
// Certified fast arithmetic for the later stages (and for stage 0's undecided windows): the same fast_weak evaluation as the
// stage-0 filter, with the stage's weights in shared memory and the per-record geometry from global memory, decides a record
// when the float sum of the fast weak outputs is on the certain side of the stage's limits (ScPlan.fl_*, built like
// ScFastParams.lim_*): certainly rejected -- and its `multi` certainly 2 or certainly 1 -- or certainly passed (not for the
// last stage, whose exact score is the detection's output).  What it cannot decide is queued per CTA and run DENSELY through
// the reference arithmetic (one queued record per thread), so a warp never executes the 690-instruction exact path for a
// single lane.  Scores of rejected records are not part of any output (k_finalize), so a certified rejection stores none.
#define SC_STAGE_Q 384   // per-CTA queue of undecided records (drained 128 at a time)
template <int HP>
__global__ void __launch_bounds__(128) k_scan_stage(const ScPlan* __restrict__ plan, int stage, const float4* __restrict__ S,
                                                     const ScGeom* __restrict__ geom_all, const float* __restrict__ w_all,
                                                     const double* __restrict__ wb_all, uint32_t* __restrict__ multi_bits,
                                                     ScRecord* __restrict__ rec, const uint32_t* __restrict__ in_idx,
                                                     const uint32_t* __restrict__ in_count, uint32_t* __restrict__ out_idx,
                                                     uint32_t* __restrict__ out_count, uint32_t cap) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ uint32_t s_q[SC_STAGE_Q];
    __shared__ uint32_t s_qn;
    const int n_weak = plan->n_weak[stage], wbase = plan->weak_base[stage], total_weak = plan->total_weak;
    float* sw = reinterpret_cast<float*>(s_dyn);
    double* swb = reinterpret_cast<double*>(s_dyn + (size_t)n_weak * SC_W_PITCH * 4);
    for (int i = threadIdx.x; i < n_weak * SC_W_PITCH; i += blockDim.x) sw[i] = w_all[(size_t)wbase * SC_W_PITCH + i];
    for (int i = threadIdx.x; i < n_weak; i += blockDim.x) swb[i] = wb_all[wbase + i];
    if (threadIdx.x == 0) s_qn = 0;
    __syncthreads();
    const uint32_t count = min(*in_count, cap);
    const int n_stages = plan->n_stages;
    constexpr int ppitch = SC_ROW_ELEMS(HP);
    const bool force = plan->force_all != 0, last = stage == n_stages - 1;
    const bool fast = plan->stage_fast != 0 && !force;
    const float theta = plan->theta[stage];
    const float lim_reject = plan->fl_reject[stage], lim_pass = plan->fl_pass[stage], lim_skip = plan->fl_skip[stage], lim_noskip = plan->fl_noskip[stage];
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t stride = gridDim.x * blockDim.x;

    auto window_base = [&](const ScRecord& r) -> const char* {
        const int f = r.fs >> 8, si = r.fs & 0xff, gy = r.yx >> 16, gx = r.yx & 0xffff;
        return reinterpret_cast<const char*>(S + (size_t)f * plan->lay.frame4 + ((gy + plan->sc[si].gy0) * ppitch + SC_COL(gx >> 1)));
    };
    auto geom_of = [&](const ScRecord& r) -> const ScGeom* {
        return geom_all + ((size_t)(r.yx & 1u) * plan->n_scales + (r.fs & 0xff)) * total_weak + wbase;
    };
    auto set_multi = [&](const ScRecord& r) {
        const int f = r.fs >> 8, si = r.fs & 0xff, gy = r.yx >> 16, gx = r.yx & 0xffff;
        atomicOr(&multi_bits[(size_t)f * plan->words_per_frame + plan->sc[si].word_base + (size_t)gy * plan->sc[si].wpr + (gx >> 5)], 1u << (gx & 31));
    };
    auto push_alive = [&](bool push, uint32_t idx) {  // warp-converged
        const uint32_t m = __ballot_sync(0xffffffffu, push);
        if (m) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(out_count, __popc(m));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (push) {
                const uint32_t slot = slot0 + __popc(m & lt);
                if (slot < cap) out_idx[slot] = idx;
            }
        }
    };
    // the reference arithmetic on one record (GentleAdaboost::Predict2, threshold, ObjDetector.cpp:196-201,214); returns "stays alive"
    auto exact_record = [&](uint32_t idx) -> bool {
        ScRecord r = rec[idx];
        const float score = stage_score<HP>(window_base(r), geom_of(r), sw, swb, n_weak, HP);
        if (r.rej >= 0) return false;  // force_all: an earlier stage already rejected it; later stages are evaluated, not recorded
        const bool rejected = score < theta;
        if (rejected) {
            r.rej = stage; r.score = __float_as_uint(score);
            if (rejected_skips(score, stage, n_stages)) set_multi(r);
            rec[idx] = r;
        } else if (last) {
            r.rej = n_stages; r.score = __float_as_uint(score);
            rec[idx] = r;
        } else if (r.rej != -1) {
            r.rej = -1;
            rec[idx] = r;
        }
        return !rejected && !last && !force;  // force_all walks the full record array at every stage
    };
    auto drain = [&](uint32_t take) {  // block-converged: the last `take` (<= 128) queue entries, one per thread
        const uint32_t qn = s_qn;
        const bool active = threadIdx.x < take;
        uint32_t idx = 0;
        bool alive = false;
        if (active) { idx = s_q[qn - take + threadIdx.x]; alive = exact_record(idx); }
        push_alive(alive, idx);
        __syncthreads();
        if (threadIdx.x == 0) s_qn = qn - take;
        __syncthreads();
    };

    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += stride) {  // block-uniform trip count
        const uint32_t i = i0 + threadIdx.x;
        bool push = false, undecided = false;
        uint32_t idx = 0;
        if (i < count) {
            idx = in_idx ? in_idx[i] : i;
            if (!fast) {
                undecided = true;
            } else {
                ScRecord r = rec[idx];
                if (stage == 0 && r.rej == -2) {          // the stage-0 filter certified the pass
                    push = true;
                } else {
                    const char* base = window_base(r);
                    const ScGeom* gq = geom_of(r);
                    float sum = 0.f;
                    for (int q = 0; q < n_weak; q++) {
                        ScGeom g;
                        const uint4* gs = reinterpret_cast<const uint4*>(gq + q);
                        const uint4 g0 = gs[0], g1 = gs[1], g2 = gs[2];
                        g.c[0] = g0.x; g.c[1] = g0.y; g.c[2] = g0.z; g.c[3] = g0.w; g.c[4] = g1.x; g.c[5] = g1.y; g.c[6] = g1.z; g.c[7] = g1.w;
                        g.c[8] = g2.x; g.c[9] = g2.y; g.shape = g2.z; g.pad = 0;
                        float v[32];
                        box_sums_p<HP>(base, g, HP, v);
                        sum = __fadd_rn(sum, fast_tail(v, sw + q * SC_W_PITCH, (float)swb[q]));
                    }
                    if (sum < lim_reject) {
                        if (sum < lim_skip) { r.rej = stage; r.score = 0u; rec[idx] = r; set_multi(r); }
                        else if (sum >= lim_noskip) { r.rej = stage; r.score = 0u; rec[idx] = r; }
                        else undecided = true;
                    } else if (sum >= lim_pass && !last) {
                        push = true;
                    } else {
                        undecided = true;
                    }
                }
            }
        }
        push_alive(push, idx);
        {   // queue what needs the reference arithmetic
            const uint32_t m = __ballot_sync(0xffffffffu, undecided);
            uint32_t q0 = 0;
            if (m && lane == 0) q0 = atomicAdd(&s_qn, __popc(m));
            q0 = __shfl_sync(0xffffffffu, q0, 0);
            if (undecided) s_q[q0 + __popc(m & lt)] = idx;   // at most 255 queued before + 128 new <= SC_STAGE_Q
        }
        __syncthreads();
        if (s_qn >= 128) drain(128);   // block-uniform (s_qn is read after the barrier by every thread)
    }
    __syncthreads();
    while (s_qn > 0) drain(min(s_qn, 128u));
}


// ---------------------------------------------------------------------------------------------------------
// Adaptive-stride replay (ObjDetector.cpp:185-186,214-217): one thread per (frame, scale, lattice row)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_replay_rows(const ScPlan* __restrict__ plan, int nframes, const uint32_t* __restrict__ multi_bits,
                                                      const uint32_t* __restrict__ pass_bits, uint32_t* __restrict__ visited_bits,
                                                      unsigned long long* __restrict__ counters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = plan->rows_per_frame;
    unsigned long long nvis = 0, npass = 0;
    int f = 0;
    if (t < nframes * rows) {
        f = t / rows;
        const int r = t - f * rows;
        int si = 0;
        while (si + 1 < plan->n_scales && plan->sc[si + 1].row_base <= r) si++;
        const int wpr = plan->sc[si].wpr, nx = plan->sc[si].nx;
        const size_t w0 = (size_t)f * plan->words_per_frame + plan->sc[si].word_base + (size_t)(r - plan->sc[si].row_base) * wpr;
        const bool skip = plan->skip_rule != 0;
        int pos = 0;
        for (int wi = 0; wi < wpr; wi++) {
            const uint32_t m = multi_bits[w0 + wi], pm = pass_bits[w0 + wi];
            const int end = min(32 * (wi + 1), nx);
            uint32_t v = 0;
            if (skip) {
                while (pos < end) {
                    const int bpos = pos & 31;
                    v |= 1u << bpos;
                    pos += 1 + ((m >> bpos) & 1u);
                }
            } else {
                const int nb = end - 32 * wi;
                v = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
            }
            visited_bits[w0 + wi] = v;
            nvis += __popc(v);
            npass += __popc(v & pm);
        }
    }
    // one atomic per warp when the whole warp works on the same frame, else one per thread
    const uint32_t same = __match_any_sync(0xffffffffu, f);
    if (same == 0xffffffffu) {
        for (int d = 16; d > 0; d >>= 1) {
            nvis += __shfl_down_sync(0xffffffffu, nvis, d);
            npass += __shfl_down_sync(0xffffffffu, npass, d);
        }
        if ((threadIdx.x & 31) == 0 && (nvis | npass)) {
            atomicAdd(&counters[(size_t)f * SC_CNT_STRIDE + SC_CNT_VISITED], nvis);
            atomicAdd(&counters[(size_t)f * SC_CNT_STRIDE + SC_CNT_PREFILTER], npass);
        }
    } else if (nvis | npass) {
        atomicAdd(&counters[(size_t)f * SC_CNT_STRIDE + SC_CNT_VISITED], nvis);
        atomicAdd(&counters[(size_t)f * SC_CNT_STRIDE + SC_CNT_PREFILTER], npass);
    }
}

// Detections = records that passed every stage AND were visited by the replay; reach counters for visited records.
struct ScDetOut { int32_t frame, x, y, l; double score; };

__global__ void __launch_bounds__(128) k_finalize(const ScPlan* __restrict__ plan, const ScRecord* __restrict__ rec,
                                                   const uint32_t* __restrict__ rec_count, uint32_t rec_cap,
                                                   const uint32_t* __restrict__ visited_bits, unsigned long long* __restrict__ counters,
                                                   ScDetOut* __restrict__ det, uint32_t* __restrict__ det_count, uint32_t det_cap, int frame0) {
    const uint32_t count = min(*rec_count, rec_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    const int n_stages = plan->n_stages;
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < count; i0 += stride) {  // warp-uniform trip count
        const uint32_t i = i0 + threadIdx.x;
        bool vis = false;
        int f = 0, gx = 0, gy = 0, ya = 0, l = 0, reached = 0;
        float score = 0.f;
        if (i < count) {
            const ScRecord r = rec[i];
            f = r.fs >> 8;
            const int si = r.fs & 0xff;
            gy = r.yx >> 16; gx = r.yx & 0xffff; l = plan->sc[si].l;
            ya = gy + plan->sc[si].gy0;  // absolute lattice row (plans of a row band count rows from gy0)
            const uint32_t v = visited_bits[(size_t)f * plan->words_per_frame + plan->sc[si].word_base + (size_t)gy * plan->sc[si].wpr + (gx >> 5)];
            vis = (v >> (gx & 31)) & 1u;
            reached = r.rej < 0 ? n_stages : r.rej;  // stages 0..min(reached, n_stages-1) were entered
            score = __uint_as_float(r.score);
        }
        unsigned long long* c = counters + (size_t)f * SC_CNT_STRIDE;
        // warp-aggregated reach counters: one atomic per (frame group, stage)
        const uint32_t grp = __match_any_sync(0xffffffffu, f);
        const bool leader = lane == (__ffs(grp) - 1);
        for (int s = 1; s < n_stages; s++) {
            const uint32_t m = __ballot_sync(0xffffffffu, vis && s <= reached) & grp;
            if (leader && m) atomicAdd(&c[SC_CNT_REACH0 + s], (unsigned long long)__popc(m));
        }
        if (vis && reached == n_stages) {
            atomicAdd(&c[SC_CNT_RAW], 1ull);
            const uint32_t slot = atomicAdd(det_count, 1u);
            if (slot < det_cap) {
                ScDetOut d;
                d.frame = frame0 + f; d.x = gx * plan->step; d.y = ya * plan->step; d.l = l;
                d.score = ((double)score + (double)n_stages + 1.0) / (double)n_stages;  // ObjDetector.cpp:201
                det[slot] = d;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Grouping of the raw windows on the device: cv::groupRectangles(wins, weights = 0.., scores, thr, eps) as called at
// ObjDetector.cpp:224-225 (OpenCV's portable implementation), one CTA per frame.
//   classes  = connected components of `similar` (partition): min-label propagation + pointer jumping; a class is named
//              by its smallest member index, so ascending names = the first-seen order of cv::partition on the
//              (l, y, x)-sorted window list the host path uses
//   per class: integer sums of x, y, w, h, member count, best score; mean = cvRound(float(sum) * (1.f / n))
//   kept     : more than thr members and not inside a larger kept-eligible class's mean rect grown by eps
// ---------------------------------------------------------------------------------------------------------
#define SC_GROUP_MAX 2048   // raw windows of one frame a CTA groups (more: the host path takes the batch)
struct ScGroupOut { int32_t frame, idx, x, y, w, h; double score; };

__global__ void k_group_count(const ScDetOut* __restrict__ det, const uint32_t* __restrict__ n_det, uint32_t cap, int frame_lo, int nframes,
                              uint32_t* __restrict__ per_frame) {
    const uint32_t n = min(*n_det, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int f = det[i].frame - frame_lo;
        if (f >= 0 && f < nframes) atomicAdd(&per_frame[f], 1u);
    }
}

// single block: exclusive offsets of the per-frame segments, fill cursors reset, overflow flag
__global__ void k_group_offsets(const uint32_t* __restrict__ per_frame, int nframes, uint32_t* __restrict__ offsets, uint32_t* __restrict__ fill,
                                uint32_t* __restrict__ flags) {
    if (threadIdx.x == 0) {
        uint32_t run = 0, over = 0;
        for (int f = 0; f < nframes; f++) { offsets[f] = run; run += per_frame[f]; fill[f] = 0; over |= per_frame[f] > SC_GROUP_MAX; }
        offsets[nframes] = run;
        flags[0] = over;
    }
}

__global__ void k_group_scatter(const ScDetOut* __restrict__ det, const uint32_t* __restrict__ n_det, uint32_t cap, int frame_lo, int nframes,
                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ fill, ScDetOut* __restrict__ seg) {
    const uint32_t n = min(*n_det, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const ScDetOut d = det[i];
        const int f = d.frame - frame_lo;
        if (f >= 0 && f < nframes) seg[offsets[f] + atomicAdd(&fill[f], 1u)] = d;
    }
}

__device__ __forceinline__ bool rects_similar(int ax, int ay, int al, int bx, int by, int bl, double eps) {
    const int m = min(al, bl);
    const double delta = eps * (double)(m + m) * 0.5;  // eps * (min(w) + min(h)) * 0.5, all windows square
    return (double)abs(ax - bx) <= delta && (double)abs(ay - by) <= delta && (double)abs(ax + al - bx - bl) <= delta &&
           (double)abs(ay + al - by - bl) <= delta;
}

__global__ void __launch_bounds__(256) k_group_frames(const ScDetOut* __restrict__ seg, const uint32_t* __restrict__ offsets, int frame_lo, int thr,
                                                       double eps, ScGroupOut* __restrict__ out, uint32_t* __restrict__ out_count, uint32_t out_cap) {
    extern __shared__ __align__(16) unsigned char g_dyn[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const uint32_t o0 = offsets[f];
    const int n = (int)(offsets[f + 1] - o0);
    if (n == 0 || n > SC_GROUP_MAX) return;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(g_dyn);      // [np2]  l << 32 | y << 16 | x ; later: best-score bits
    double* score = reinterpret_cast<double*>(key + SC_GROUP_MAX);               // [n]
    int* pay = reinterpret_cast<int*>(score + SC_GROUP_MAX);                     // [np2]  index into seg; later: class member count
    int* label = pay + SC_GROUP_MAX;                                             // [n]
    int* sx = label + SC_GROUP_MAX;                                              // [n]  class sums, indexed by the class name (a member index)
    int* sy = sx + SC_GROUP_MAX;
    int* sl = sy + SC_GROUP_MAX;
    int* roots = sl + SC_GROUP_MAX;                                              // [n]  class names in ascending order; then kept classes
    __shared__ int s_nroots;
    __shared__ uint32_t s_base;
    // ---- sort by (l, y, x): the order the host path groups in
    for (int i = tid; i < np2; i += 256) {
        if (i < n) {
            const ScDetOut d = seg[o0 + i];
            key[i] = ((unsigned long long)(uint32_t)d.l << 32) | ((unsigned long long)(uint32_t)d.y << 16) | (unsigned long long)(uint32_t)d.x;
            pay[i] = i;
        } else {
            key[i] = ~0ull; pay[i] = -1;
        }
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += 256) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[p];
                    if ((a > b) == up) { key[i] = b; key[p] = a; const int t = pay[i]; pay[i] = pay[p]; pay[p] = t; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += 256) { score[i] = seg[o0 + pay[i]].score; label[i] = i; }
    __syncthreads();
    // ---- connected components of `similar`: every window takes the smallest label among its similar windows, then jumps
    for (;;) {
        int changed = 0;
        for (int i = tid; i < n; i += 256) {
            const unsigned long long ki = key[i];
            const int xi = (int)(ki & 0xffffu), yi = (int)((ki >> 16) & 0xffffu), li = (int)(ki >> 32);
            int m = label[i];
            // similar windows differ in side by at most 2 delta = 2 eps min(l): in the (l, y, x)-sorted list the candidates
            // are the contiguous run with l in [li / (1 + 2 eps), li (1 + 2 eps)] (found by bisection; a superset is fine)
            const double span = 1.0 + 2.0 * eps;
            const unsigned long long klo = (unsigned long long)(uint32_t)max((int)floor((double)li / span) - 1, 0) << 32;
            const unsigned long long khi = ((unsigned long long)(uint32_t)((int)ceil((double)li * span) + 1) << 32) | 0xffffffffull;
            int j0 = 0, j1 = n;
            for (int a = 0, b = n; a < b;) { const int mid = (a + b) >> 1; if (key[mid] < klo) a = mid + 1; else b = mid; j0 = a; }
            for (int a = j0, b = n; a < b;) { const int mid = (a + b) >> 1; if (key[mid] <= khi) a = mid + 1; else b = mid; j1 = a; }
            for (int j = j0; j < j1; j++) {
                const int lj = label[j];
                if (lj < m) {
                    const unsigned long long kj = key[j];
                    if (rects_similar(xi, yi, li, (int)(kj & 0xffffu), (int)((kj >> 16) & 0xffffu), (int)(kj >> 32), eps)) m = lj;
                }
            }
            if (m < label[i]) { label[i] = m; changed = 1; }
        }
        __syncthreads();
        for (int i = tid; i < n; i += 256) {
            int r = label[i];
            while (label[r] != r) r = label[r];
            label[i] = r;   // racing writers only ever lower a label towards its root
        }
        if (!__syncthreads_or(changed)) break;
    }
    // ---- per class: sums, member count, best score
    for (int i = tid; i < n; i += 256) { sx[i] = 0; sy[i] = 0; sl[i] = 0; pay[i] = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const unsigned long long ki = key[i];
        const int r = label[i];
        atomicAdd(&sx[r], (int)(ki & 0xffffu)); atomicAdd(&sy[r], (int)((ki >> 16) & 0xffffu)); atomicAdd(&sl[r], (int)(ki >> 32));
        atomicAdd(&pay[r], 1);
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) key[i] = (unsigned long long)__double_as_longlong(DBL_MIN);  // keys are consumed: reuse as best-score bits
    __syncthreads();
    for (int i = tid; i < n; i += 256)
        if (score[i] > DBL_MIN) atomicMax(&key[label[i]], (unsigned long long)__double_as_longlong(score[i]));  // positive doubles order like their bits
    __syncthreads();
    // ---- class list in ascending name order (= first-seen order), mean rects (into sx / sy / sl)
    if (tid < 32) {
        int run = 0;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + tid;
            const bool is_root = i < n && label[i] == i;
            const uint32_t m = __ballot_sync(0xffffffffu, is_root);
            if (is_root) roots[run + __popc(m & ((1u << tid) - 1u))] = i;
            run += __popc(m);
        }
        if (tid == 0) s_nroots = run;
    }
    __syncthreads();
    const int k_cls = s_nroots;
    for (int c = tid; c < k_cls; c += 256) {
        const int r = roots[c];
        const float inv = __fdiv_rn(1.f, (float)pay[r]);
        sx[r] = __float2int_rn(__fmul_rn((float)sx[r], inv));
        sy[r] = __float2int_rn(__fmul_rn((float)sy[r], inv));
        sl[r] = __float2int_rn(__fmul_rn((float)sl[r], inv));
    }
    __syncthreads();
    // ---- keep: > thr members and not swallowed; label[] is free now: kept flag per class slot
    for (int c = tid; c < k_cls; c += 256) {
        const int r = roots[c], ni = pay[r];
        bool keep = ni > thr;
        if (keep) {
            const int ax = sx[r], ay = sy[r], al = sl[r];
            for (int d = 0; d < k_cls && keep; d++) {
                const int q = roots[d], nj = pay[q];
                if (d == c || nj <= thr) continue;
                const int bx = sx[q], by = sy[q], bl = sl[q];
                const int dd = __double2int_rn((double)bl * eps);
                if (ax >= bx - dd && ay >= by - dd && ax + al <= bx + bl + dd && ay + al <= by + bl + dd && (nj > max(3, ni) || ni < 3)) keep = false;
            }
        }
        label[c] = keep ? 1 : 0;
    }
    __syncthreads();
    if (tid < 32) {
        int run = 0;
        for (int b0 = 0; b0 < k_cls; b0 += 32) {
            const int c = b0 + tid;
            const bool kp = c < k_cls && label[c] != 0;
            const uint32_t m = __ballot_sync(0xffffffffu, kp);
            if (kp) label[c] = 1 + run + __popc(m & ((1u << tid) - 1u));   // 1-based output rank
            run += __popc(m);
        }
        if (tid == 0) s_base = run ? atomicAdd(out_count, (uint32_t)run) : 0u;
    }
    __syncthreads();
    for (int c = tid; c < k_cls; c += 256) {
        if (label[c] == 0) continue;
        const int r = roots[c];
        const uint32_t slot = s_base + (uint32_t)(label[c] - 1);
        if (slot < out_cap) {
            ScGroupOut g;
            g.frame = frame_lo + f; g.idx = label[c] - 1; g.x = sx[r]; g.y = sy[r]; g.w = sl[r]; g.h = sl[r];
            g.score = __longlong_as_double((long long)key[r]);
            out[slot] = g;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Parity hooks on explicit rect / window lists (layout step 1)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_features(const float4* __restrict__ S, const ScLayout L, const int4* __restrict__ rects, int n, float* __restrict__ out,
                           float* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 r = rects[i];  // x, y, w, h
    const int pitch = L.ppitch;
    const char* base = reinterpret_cast<const char*>(S + ((size_t)r.y * pitch + SC_COL(r.x)));
    if (sums) {
        const uint32_t pf[4] = {0u, 16u * SC_COL(r.z), 16u * (uint32_t)(r.w * pitch), 16u * (uint32_t)(r.w * pitch + SC_COL(r.z))};
        sums[i] = window_sum(base, pf);
    }
    if (out) {
        ScGeom g;
        g.pad = 0;
        if (r.z == r.w) {
            const int ce = r.z / 2;
            g.shape = 0;
            for (int b = 0; b < 3; b++)
                for (int a = 0; a < 3; a++) g.c[3 * b + a] = 16u * (uint32_t)(b * ce * pitch + SC_COL(a * ce));
            g.c[9] = 0;
        } else {
            const int ce = min(r.z, r.w);
            const int along = r.z > r.w ? SC_COL(ce) : ce * pitch, across = r.z > r.w ? ce * pitch : SC_COL(ce);
            g.shape = 1;
            for (int k = 0; k < 5; k++) { g.c[k] = 16u * (uint32_t)(k * along); g.c[5 + k] = 16u * (uint32_t)(k * along + across); }
        }
        float v[32];
        descriptor<0>(base, g, L.hp, v);
        for (int k = 0; k < 32; k++) out[(size_t)i * 32 + k] = v[k];
    }
}

// Parity hook: CalcFeature's box sums (no Normalize) of explicit rects from the compact plane (layout step 1).
__global__ void k_box_sums_compact(const float4* __restrict__ S, const ScLayout L, const int4* __restrict__ rects, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 r = rects[i];  // x, y, w, h
    const int pitch = L.ppitch;
    const char* base = reinterpret_cast<const char*>(S + ((size_t)r.y * pitch + SC_COL(r.x)));
    ScGeom g;
    g.pad = 0;
    if (r.z == r.w) {
        const int ce = r.z / 2;
        g.shape = 0;
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < 3; a++) g.c[3 * b + a] = 16u * (uint32_t)(b * ce * pitch + SC_COL(a * ce));
        g.c[9] = 0;
    } else {
        const int ce = min(r.z, r.w);
        const int along = r.z > r.w ? SC_COL(ce) : ce * pitch, across = r.z > r.w ? ce * pitch : SC_COL(ce);
        g.shape = 1;
        for (int k = 0; k < 5; k++) { g.c[k] = 16u * (uint32_t)(k * along); g.c[5 + k] = 16u * (uint32_t)(k * along + across); }
    }
    float v[32];
    box_sums_c<0>(base, g, L.hp, v);
    for (int k = 0; k < 32; k++) out[(size_t)i * 32 + k] = v[k];
}

// Hard-negative mining (next row N2): the descriptors of ALL pool patches projected into each accepted window
// (FillNegSamples, DenseSURFFeatureExtractor.cpp:157-176), read from the integral images the mining scan itself produced
// (its layout L, frame slots of the scan batch) -- no second IntegralImage per image, no host-built rect list.
// One thread per (window, pool patch).  ProjectPatches' float arithmetic (:486-508) is restated on the device operation
// for operation: scale = fl(l / tmpl), x' = (int)fl(x * scale), the short side (int)fl(side * scale), the long side its multiple.
struct ScMineWin { int32_t slot, x, y, l; };
__global__ void __launch_bounds__(128) k_mine_descriptors(const float4* __restrict__ S, const ScLayout L, const ScMineWin* __restrict__ wins, int n_wins,
                                                           const int4* __restrict__ pool, int P, int tmpl, float* __restrict__ X) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_wins * P) return;
    const int k = (int)(t / P), pi = (int)(t - (long long)k * P);
    const ScMineWin w = wins[k];
    const int4 pt = pool[pi];  // x, y, w, h on the template
    const float scale = __fdiv_rn((float)w.l, (float)tmpl);
    int rx = (int)__fmul_rn((float)pt.x, scale) + w.x, ry = (int)__fmul_rn((float)pt.y, scale) + w.y, rw, rh;
    if (pt.z >= pt.w) { rh = (int)__fmul_rn((float)pt.w, scale); rw = rh * (pt.z / pt.w); }
    else { rw = (int)__fmul_rn((float)pt.z, scale); rh = rw * (pt.w / pt.z); }
    ScGeom g;
    g.pad = 0;
    if (rw == rh) {
        const int ce = rw / 2;
        g.shape = 0;
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < 3; a++) g.c[3 * b + a] = (uint32_t)(16 * sc_layout_index(L, rx + a * ce, ry + b * ce));
        g.c[9] = 0;
    } else {
        const int ce = min(rw, rh);
        const bool wide = rw > rh;
        g.shape = 1;
        for (int c = 0; c < 5; c++) {
            const int x0 = rx + (wide ? c * ce : 0), y0 = ry + (wide ? 0 : c * ce);
            g.c[c] = (uint32_t)(16 * sc_layout_index(L, x0, y0));
            g.c[5 + c] = (uint32_t)(16 * sc_layout_index(L, x0 + (wide ? 0 : ce), y0 + (wide ? ce : 0)));
        }
    }
    float v[32];
    descriptor<0>(reinterpret_cast<const char*>(S + (size_t)w.slot * L.frame4), g, L.hp, v);
    float4* o = reinterpret_cast<float4*>(X + ((size_t)k * P + pi) * 32);
#pragma unroll
    for (int c = 0; c < 8; c++) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// Training-side descriptor extraction (next row N3): ExtractNextImageFeatures -> IntegralImage + ExtractFeatures over the
// template pool (DenseSURFFeatureExtractor.cpp:65-120) for a batch of template-sized samples.  One CTA per sample, nothing
// but the sample and its descriptors touches HBM: channels, exact integer row prefixes and the reference's sequential
// float32 column recurrence are built in shared memory ((T+1)^2 pixels x 32 B; 54 KB for T = 40), then every thread takes
// pool patches and runs the exact CalcFeature + Normalize on shared-memory corners.  X[n][p][0..31].
__global__ void __launch_bounds__(320) k_pool_features(const uint8_t* __restrict__ imgs, int T, const int4* __restrict__ pool, int P, float* __restrict__ X) {
    extern __shared__ __align__(16) unsigned char f_dyn[];
    float* S = reinterpret_cast<float*>(f_dyn);                     // [(T+1)][(T+1)][8]
    int* Si = reinterpret_cast<int*>(f_dyn);
    uint8_t* im = f_dyn + (size_t)(T + 1) * (T + 1) * 32;           // [T][T]
    const int n = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, T1 = T + 1;
    const uint8_t* src = imgs + (size_t)n * T * T;
    for (int i = tid; i < T * T; i += nt) im[i] = src[i];
    for (int i = tid; i < T1 * 8; i += nt) { S[i] = 0.f; S[(size_t)(i >> 3) * T1 * 8 + (i & 7)] = 0.f; }  // row 0 and column 0
    __syncthreads();
    // exact row prefixes of the eight channels (T2bFilter :199-349): thread = (row, channel); the four differences are
    // one formula with per-channel neighbour offsets (no divergence between the channels of a warp)
    for (int t = tid; t < T * 8; t += nt) {
        const int y = t >> 3, c = t & 7, k = c >> 1;
        // d = I[y + ay][x + ax] - I[y + by][x + bx]:  dx (0,+1)-(0,-1)  dy (+1,0)-(-1,0)  du (+1,+1)-(-1,-1)  dv (-1,+1)-(+1,-1)
        const int ay = k == 0 ? 0 : (k == 3 ? -1 : 1), ax = k == 1 ? 0 : 1;
        const int ra = min(max(y + ay, 0), T - 1) * T, rb = min(max(y - ay, 0), T - 1) * T;
        const int sgn = (c & 1) ? 1 : -1;
        int run = 0;
        int* out = Si + ((size_t)(y + 1) * T1 + 1) * 8 + c;
        for (int x = 0; x < T; x++) {
            const int xa = min(max(x + ax, 0), T - 1), xb = min(max(x - ax, 0), T - 1);
            const int d = (int)im[ra + xa] - (int)im[rb + xb];
            run += max(sgn * d, 0);
            out[x * 8] = run;
        }
    }
    __syncthreads();
    // S[y+1][x+1][c] = fl32(S[y][x+1][c] + float(rowprefix)), sequential in y (cv::integral, Appendix A.2): thread = (column, channel)
    for (int t = tid; t < T * 8; t += nt) {
        const int x = (t >> 3) + 1, c = t & 7;
        float acc = 0.f;
        for (int y = 1; y <= T; y++) {
            const size_t i = ((size_t)y * T1 + x) * 8 + c;
            acc = __fadd_rn(acc, (float)Si[i]);
            S[i] = acc;
        }
    }
    __syncthreads();
    const int pitch = T1 * 32;  // bytes per integral row
    for (int p = tid; p < P; p += nt) {
        const int4 r = pool[p];  // x, y, w, h on the template
        ScGeom g;
        g.pad = 0;
        if (r.z == r.w) {
            const int ce = r.z / 2;
            g.shape = 0;
            for (int b = 0; b < 3; b++)
                for (int a = 0; a < 3; a++) g.c[3 * b + a] = (uint32_t)(b * ce * pitch + a * ce * 32);
            g.c[9] = 0;
        } else {
            const int ce = min(r.z, r.w);
            const int along = r.z > r.w ? ce * 32 : ce * pitch, across = r.z > r.w ? ce * pitch : ce * 32;
            g.shape = 1;
            for (int k = 0; k < 5; k++) { g.c[k] = (uint32_t)(k * along); g.c[5 + k] = (uint32_t)(k * along + across); }
        }
        float v[32];
        descriptor<-1>(reinterpret_cast<const char*>(S) + (size_t)r.y * pitch + (size_t)r.x * 32, g, 1, v);
        float4* o = reinterpret_cast<float4*>(X + ((size_t)n * P + p) * 32);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
}

__global__ void k_stage_scores(const ScPlan* __restrict__ plan, const float4* __restrict__ S, const ScGeom* __restrict__ geom_win,
                               const float* __restrict__ w_all, const double* __restrict__ wb_all, const int* __restrict__ wins, int n,
                               float* __restrict__ out) {
    // geom_win: [n][total_weak] geometry projected for each explicit window's side (host-built, layout step 1)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char* base = reinterpret_cast<const char*>(S + ((size_t)wins[3 * i + 1] * plan->lay.ppitch + SC_COL(wins[3 * i])));
    for (int s = 0; s < plan->n_stages; s++) {
        const int wb = plan->weak_base[s];
        out[(size_t)i * plan->n_stages + s] = stage_score<0>(base, geom_win + (size_t)i * plan->total_weak + wb, w_all + (size_t)wb * SC_W_PITCH,
                                                             wb_all + wb, plan->n_weak[s], plan->lay.hp);
    }
}

// Parity hook of the stage-0 fast filter: for explicit windows, the float sum of the fast weak outputs and the float sum
// of the reference-arithmetic weak outputs (GentleAdaboost::Predict2's accumulator before the division).
__global__ void k_stage0_fast_check(const ScPlan* __restrict__ plan, const float4* __restrict__ S, const ScGeom* __restrict__ geom_win,
                                    const float* __restrict__ w_all, const double* __restrict__ wb_all, const int* __restrict__ wins, int n,
                                    float* __restrict__ fast_sum, float* __restrict__ exact_sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char* base = reinterpret_cast<const char*>(S + ((size_t)wins[3 * i + 1] * plan->lay.ppitch + SC_COL(wins[3 * i])));
    float fs = 0.f, es = 0.f;
    for (int q = 0; q < plan->n_weak[0]; q++) {
        const ScGeom g = geom_win[(size_t)i * plan->total_weak + q];
        float wl[32];
#pragma unroll
        for (int k = 0; k < 32; k++) wl[k] = w_all[(size_t)q * SC_W_PITCH + k];
        float v[32];
        box_sums_p<0>(base, g, plan->lay.hp, v);
        fs = __fadd_rn(fs, fast_tail(v, wl, (float)wb_all[q]));
        descriptor<0>(base, g, plan->lay.hp, v);
        es = __fadd_rn(es, weak_predict(v, w_all + (size_t)q * SC_W_PITCH, wb_all[q]));
    }
    fast_sum[i] = fs; exact_sum[i] = es;
}

// LogisticRegression::Predict on explicit (weights, descriptor) pairs; with `mean` non-null thread 0 also folds
// the probabilities in order into GentleAdaboost::Predict2's float32 mean.
__global__ void k_weak_predict(const float* __restrict__ w36, const double* __restrict__ wb, const float* __restrict__ x, int n,
                               float* __restrict__ out, float* __restrict__ mean) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = x[(size_t)i * 32 + k];
        out[i] = weak_predict(v, w36 + (size_t)i * SC_W_PITCH, wb[i]);
    }
    if (mean) {  // single-block launch
        __syncthreads();
        if (threadIdx.x == 0) {
            float acc = 0.f;
            for (int i = 0; i < n; i++) acc = __fadd_rn(acc, out[i]);
            *mean = __fdiv_rn(acc, (float)n);
        }
    }
}

// Measurement probe (no product role): every thread gathers `per_thread` random 32-byte sectors (two 16-byte loads, as
// a corner fetch does) from a table of n_sectors sectors; establishes the L2 / HBM sector-gather ceiling that
// MEASURED_PEAKS.json lacks.
__global__ void __launch_bounds__(256) k_probe_gather(const float4* __restrict__ table, uint32_t n_sectors, int per_thread, float* __restrict__ sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int i = 0; i < per_thread; i += 4) {
        uint32_t idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { s = s * 1664525u + 1013904223u; idx[k] = (uint32_t)(((unsigned long long)(s >> 4) * n_sectors) >> 28); }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 a = __ldg(table + 2 * (size_t)idx[k]), b = __ldg(table + 2 * (size_t)idx[k] + 1);
            acc += a.x + b.w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;  // keep the loads alive
}

// Measurement probe (no product role): coalesced 16-byte loads (512 contiguous bytes per warp, as one corner fetch of the
// scan in the half-split layout) streaming over an L2-resident table; mode 0 = ld.global.cg (L2 only), 1 = ld.global.nc
// (allocates in L1).  Establishes the L2 -> SM bandwidth ceiling the scan's sector traffic is set against.
template <int U>
__global__ void __launch_bounds__(256) k_probe_stream(const float4* __restrict__ table, uint32_t n4_mask, int per_thread, int mode, float* __restrict__ sink) {
    // table length is a power of two (n4_mask = length - 1): every CTA streams its own contiguous window, wrapping
    float acc = 0.f;
    const uint32_t pos = blockIdx.x * (uint32_t)per_thread * 256u + threadIdx.x;
    for (int i = 0; i < per_thread; i += U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const uint32_t p = (pos + (uint32_t)(i + k) * 256u) & n4_mask;
            v[k] = mode ? __ldg(table + p) : __ldcg(table + p);
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc += v[k].x + v[k].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// Training-side pool evaluation: candidate scoring of one boosting round,
// GentleAdaboost.cpp:145-148 -> StageClassifier::Evaluate (StageClassifier.cpp:35-70) over
// GentleAdaboost::Predict (GentleAdaboost.cpp:233-245).  HBM-bound stream over X [N][P][32].
// ---------------------------------------------------------------------------------------------------------
#define SC_POOL_LEVELS 21   // 20 thresholds 1, 1-0.05f, ... (float t = 1; t >= 0; t -= 0.05f) + "below all"
#define SC_POOL_KC 32       // candidates per CTA tile
#define SC_POOL_NC 8        // samples per CTA tile

// hist[k][cls][level]: number of samples of class cls (1 = positive) whose stage probability with candidate k first
// reaches threshold index `level` (level 20 = below every threshold)
__global__ void __launch_bounds__(256) k_pool_hist(const float* __restrict__ X, int N, int P, const uint8_t* __restrict__ labels,
                                                    const float* __restrict__ w36, const double* __restrict__ wb,
                                                    const float* __restrict__ prior_sum, float inv_count_as_divisor,
                                                    const float* __restrict__ thresholds, int sample_slices, uint32_t* __restrict__ hist) {
    __shared__ float s_x[SC_POOL_NC * SC_POOL_KC][33];
    __shared__ float s_w[SC_POOL_KC][33];
    __shared__ double s_wb[SC_POOL_KC];
    __shared__ float s_thr[SC_POOL_LEVELS - 1];
    __shared__ uint32_t s_hist[SC_POOL_KC][2][SC_POOL_LEVELS];
    const int tid = threadIdx.x;
    const int kc = blockIdx.x / sample_slices, slice = blockIdx.x - kc * sample_slices;
    const int k0 = kc * SC_POOL_KC;
    for (int i = tid; i < SC_POOL_KC * 33; i += 256) {
        const int k = i / 33, j = i - k * 33;
        s_w[k][j] = (k0 + k < P) ? w36[(size_t)(k0 + k) * SC_W_PITCH + j] : 0.f;
    }
    if (tid < SC_POOL_KC) s_wb[tid] = (k0 + tid < P) ? wb[k0 + tid] : 0.0;
    if (tid < SC_POOL_LEVELS - 1) s_thr[tid] = thresholds[tid];
    for (int i = tid; i < SC_POOL_KC * 2 * SC_POOL_LEVELS; i += 256) (&s_hist[0][0][0])[i] = 0;
    __syncthreads();
    const int kl = tid & 31, nl = tid >> 5;
    const int n_chunks = (N + SC_POOL_NC - 1) / SC_POOL_NC;
    for (int chunk = slice; chunk < n_chunks; chunk += sample_slices) {
        const int n0 = chunk * SC_POOL_NC;
        // coalesced tile load: per sample 32 candidates x 128 B contiguous
#pragma unroll
        for (int it = 0; it < SC_POOL_NC; it++) {
            const int n = n0 + it;
            const int e = tid;                       // float4 index inside the 4 KB row: candidate e / 8, quad e % 8
            const int k = e >> 3, q = e & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N && k0 + k < P) v = __ldg(reinterpret_cast<const float4*>(X + ((size_t)n * P + k0 + k) * 32) + q);
            float* dst = &s_x[it * SC_POOL_KC + k][4 * q];
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        __syncthreads();
        const int n = n0 + nl;
        if (n < N && k0 + kl < P) {
            // LogisticRegression::Predict, same operation order as weak_predict()
            const float* x = s_x[nl * SC_POOL_KC + kl];
            const float* w = s_w[kl];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                s0 = __fadd_rn(__fmul_rn(w[i], x[i]), s0);
                s1 = __fadd_rn(__fmul_rn(w[i + 1], x[i + 1]), s1);
                s2 = __fadd_rn(__fmul_rn(w[i + 2], x[i + 2]), s2);
                s3 = __fadd_rn(__fmul_rn(w[i + 3], x[i + 3]), s3);
            }
            const float z32 = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
            const double z = (double)z32 + s_wb[kl];
            const float p = (float)(1.0 / (1.0 + exp(-z)));
            // GentleAdaboost::Predict: float running sum (prior + candidate last), divided by the classifier count
            const float prob = __fdiv_rn(__fadd_rn(prior_sum ? prior_sum[n] : 0.f, p), inv_count_as_divisor);
            int level = 0;
#pragma unroll
            for (int i = 0; i < SC_POOL_LEVELS - 1; i++) level += !(prob >= s_thr[i]);  // thresholds decrease
            atomicAdd(&s_hist[kl][labels[n] ? 1 : 0][level], 1u);
        }
        __syncthreads();
    }
    for (int i = tid; i < SC_POOL_KC * 2 * SC_POOL_LEVELS; i += 256) {
        const int k = i / (2 * SC_POOL_LEVELS);
        const uint32_t v = (&s_hist[0][0][0])[i];
        if (v && k0 + k < P) atomicAdd(&hist[(size_t)k0 * 2 * SC_POOL_LEVELS + i], v);
    }
}

// StageClassifier::Evaluate's trapezoid AUC from the level histograms (float arithmetic as StageClassifier.cpp:59-66)
__global__ void k_pool_auc(const uint32_t* __restrict__ hist, int P, float n_pos, float n_neg, float* __restrict__ auc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t* hn = hist + (size_t)k * 2 * SC_POOL_LEVELS;  // class 0 = negatives
    const uint32_t* hp = hn + SC_POOL_LEVELS;
    unsigned long long cp = 0, cn = 0;
    float area = 0.f, tpr_prev = 0.f, fpr_prev = 0.f;
    for (int i = 0; i < SC_POOL_LEVELS - 1; i++) {
        cp += hp[i]; cn += hn[i];  // samples with prob >= t_i
        const float tpr = __fdiv_rn((float)cp, n_pos), fpr = __fdiv_rn((float)cn, n_neg);
        if (i > 0) area = __fadd_rn(area, __fmul_rn(__fmul_rn(__fadd_rn(tpr, tpr_prev), __fsub_rn(fpr, fpr_prev)), 0.5f));
        tpr_prev = tpr; fpr_prev = fpr;
    }
    auc[k] = area;
}

}  // namespace sck

#endif
