// Host/device-shared plan structures for the SURF-cascade scan (sm_100a).
#ifndef SC_PLAN_H
#define SC_PLAN_H

#include <stdint.h>

#define SC_PLAN_MAX_SCALES 64
#define SC_PLAN_MAX_STAGES 16
#define SC_W_PITCH 36          // floats per weak classifier in device memory: w[0..32] + 3 pad (16 B aligned rows)

// Stage-0 tile: 64 x 16 window origins of ONE x-parity per CTA (lattice columns 2j + phase, j in a 64-run: four
// 32-wide bitmask words per tile row).  The reference's adaptive stride (multi = 2 after almost every window) walks
// the even columns until the first window that does not skip; the scan therefore evaluates the even columns first
// (phase 0) and the odd columns only from that window on (phase 1), about 56-58 % of the lattice at 1080p.
#define SC_TILE_X 64
// Tile rows: 32 for the fast-filter variant (measured on C2: 8 -> 0.270, 16 -> 0.2525, 32 -> 0.2477, 64 -> 0.325 ms/frame for
// the even columns), 16 for the exact variant (force_all / more than SC_F_MAXW weak classifiers in stage 0; on C4 the
// record order of 32-row tiles costs the later stages 7 %).
#ifndef SC_TILE_Y_FAST
#define SC_TILE_Y_FAST 32
#endif
#ifndef SC_TILE_Y_EXACT
#define SC_TILE_Y_EXACT 16
#endif
#ifndef SC_TILE_THREADS
#define SC_TILE_THREADS 256   // tile rows must be a multiple of the warp count; 4 * rows <= SC_TILE_THREADS
#endif
#ifndef SC_STAGE0_MIN_CTAS
#define SC_STAGE0_MIN_CTAS 3   // 80 registers: 3 CTAs (24 warps) per SM; 4 forces 64 registers and spills (measured slower)
#endif

#ifndef SC_FAST_MIN_CTAS
#define SC_FAST_MIN_CTAS 4     // the fast-filter kernel holds no exact arithmetic: 64 registers, 4 CTAs (32 warps) per SM
#endif

// k_scan_odd work unit: a run of up to SC_ODD_UNIT reachable odd-column windows of one lattice row (one warp per unit:
// prefilter in four 32-window passes, the survivors compacted into a dense list for the fast filter).
#ifndef SC_ODD_UNIT
#define SC_ODD_UNIT 128
#endif

// Integral strips: one warp walks one 32-column strip down the frame, SC_WALK_RB rows per prefetch block.
#define SC_STRIP 32
#ifndef SC_WALK_TILED
#define SC_WALK_TILED 1        // k_integral_walk_tiled (row prefixes by lanes turned 90 degrees) instead of k_integral_walk (shuffle scans)
#endif
#ifndef SC_WALK_RB
#define SC_WALK_RB 4
#endif

// Integral-image layout in HBM ("lattice-deinterleaved, half-split, with a compact exact-integer third"):
//   the reference keeps 8 interleaved floats (32 B) per pixel; a warp of 32 neighbouring windows on a step-s
//   lattice would then touch 32 sectors in 16 cache lines per 16-byte load.  Here pixel (X, Y) lives in plane
//   (Y mod s, X mod s) at (Y div s, X div s), and every plane row is stored as three runs of hp 16-byte elements:
//       [hp float4: channels 0-3][hp float4: channels 4-7][hp uint4: compact integer integral N]
//   Neighbouring lattice windows read neighbouring 16-byte elements: one corner fetch of a warp is fully coalesced
//   512-byte loads.  Detection plans deinterleave rows by sy = lattice step and columns by sx = 2 * step, because one
//   launch of the scan walks lattice columns of a single parity (2j or 2j + 1, see SC_TILE_X): consecutive j are then
//   consecutive elements.  The explicit-rect hooks use sx = sy = 1.  hp is a power of two, so the second and third
//   run of a pixel sit at the first one's address plus a compile-time constant (the scan kernels are instantiated
//   per hp) and cost no address arithmetic.
//
//   Compact plane: word k of N holds  (I[2k+1] << 16) + I[2k]  mod 2^32, with I[c] the EXACT integer integral of channel c
//   (the walk kernels carry it beside the float32 recurrence).  N is linear in the integrals, so for a box
//       N(A) + N(D) - N(B) - N(C)  =  (V[2k+1] << 16) + V[2k]   (mod 2^32),   V[c] = exact box sum of channel c;
//   whenever every V[c] < 65536 the two 16-bit fields ARE the box sums, and they equal the reference's float32 box
//   sums bit for bit as long as the integrals involved are below 2^23 (then fl(A + D), fl(B + C) and their difference are
//   exact).  The stage-0 fast filter reads 16 bytes per corner instead of 32 wherever both conditions are certified
//   (k_cell_bounds + the tile's far-corner check); the scan is bound by L2 -> SM bytes, so this is where its time goes.
#define SC_COL(x) (x)                 // 16-byte element index of plane column x inside a plane row
#define SC_HI(hp) (hp)                // element distance from a pixel's channels 0-3 to its channels 4-7
#define SC_NOFF(hp) (2 * (hp))        // element distance from a pixel's channels 0-3 to its compact integer word
#ifdef SC_EXP_LAYOUT2   // timing experiment only: round-1 row pitch (the compact plane then aliases the next row: results wrong)
#define SC_ROW_ELEMS(hp) (2 * (hp))
#else
#define SC_ROW_ELEMS(hp) (3 * (hp))   // 16-byte elements per plane row
#endif
#define SC_CELL_LIMIT 65536u          // a compact box sum is valid below this
struct ScLayout {
    int sx, sy;            // column / row deinterleave factors
    int hp;                // 16-byte elements per run of a plane row: power of two >= ceil((W+1)/sx), one of 256..4096
    int ppitch;            // 16-byte elements per plane row  = 3 * hp (SC_ROW_ELEMS)
    int prows;             // plane rows                      = ceil((H+1)/sy)
    int pad;
    long long plane4;      // float4 elements per plane       = prows * ppitch
    long long frame4;      // float4 elements per frame       = sx * sy * plane4
};

struct ScScale {
    int l;            // window side (int)(base * scale^i), ObjDetector.cpp:180
    int nx, ny;       // window origins per row / rows on the step lattice
    int wpr;          // 32-bit bitmask words per lattice row = ceil(nx / 32)
    float thr;        // float(l * l * prefilter), ObjDetector.cpp:188
    int tiles_x;      // stage-0 tiles per row of tiles
    int block_base;   // first stage-0 CTA of this scale inside a frame
    int word_base;    // first bitmask word of this scale inside a frame
    int row_base;     // first lattice row of this scale inside a frame (replay threads)
    uint32_t pf[2][4];  // per column parity: byte offsets of the prefilter corners (0,0) (l,0) (0,l) (l,l) relative to
                        // the window's layout element gy * ppitch + (gx >> 1)
    int gy0;          // first lattice row of this scale that the plan scans (row band of a multi-GPU split; else 0);
                      // ny, the bitmasks and the records count rows from gy0
    int pad[2];
};

struct ScPlan {
    int W, H;
    int step;
    int n_scales;
    int n_stages;
    int total_weak;
    int use_prefilter, skip_rule, force_all;
    int blocks_per_frame;         // stage-0 CTAs per frame
    int words_per_frame;          // bitmask words per frame
    int rows_per_frame;           // lattice rows per frame (all scales)
    int n_strips;                 // ceil(W / 32)
    int pad0;
    long long windows_per_frame;  // grid windows per frame
    ScLayout lay;
    float theta[SC_PLAN_MAX_STAGES];
    // certified fast arithmetic of k_scan_stage, per stage, on the float sum of the stage's fast weak outputs (as ScFastParams.lim_*)
    int stage_fast, pad1[3];
    float fl_reject[SC_PLAN_MAX_STAGES], fl_pass[SC_PLAN_MAX_STAGES], fl_skip[SC_PLAN_MAX_STAGES], fl_noskip[SC_PLAN_MAX_STAGES];
    int n_weak[SC_PLAN_MAX_STAGES];
    int weak_base[SC_PLAN_MAX_STAGES];
    ScScale sc[SC_PLAN_MAX_SCALES];
};

// Per (column parity, scale, weak classifier) projected geometry: layout offsets (float4 units, low half) of the
// patch's corner lattice relative to the window's layout index gy * ppitch + (gx >> 1)  [hooks: y * ppitch + x].
//   shape 0, square 2x2 cells: c[3*b + a], a, b in 0..2   (corner (ox + a*ce, oy + b*ce))
//   shape 1, long 4x1 / 1x4  : c[k] first line, c[5 + k] second line, k in 0..4 along the cell chain
// Offsets are BYTES (float4 index * 16), unsigned 32-bit: frames up to 4 GiB of integral image.
struct ScGeom {
    uint32_t c[10];
    int shape;
    int pad;
};

// Stage-0 fast filter (k_scan_stage0f): everything one launch needs, passed BY VALUE as a __grid_constant__ kernel
// parameter so that weights and corner offsets are read through the constant bank into uniform registers (no shared-
// memory wavefronts, no per-thread registers).  Used when the stage has at most SC_F_MAXW weak classifiers.
#define SC_F_MAXW 6
struct ScFastParams {
    int n_weak, n_scales, blocks_per_frame, pad0;
    // certified decisions on the float sum of the fast weak outputs (see fast_weak() for the error budget):
    //   sum <  lim_reject                  -> the reference's stage score is < theta for certain
    //   sum <  lim_skip / >= lim_noskip    -> the rejected window's `multi` is 2 / 1 for certain
    //   sum >= lim_pass                    -> the stage score is >= theta for certain (no exact stage 0 needed; +inf when the
    //                                         cascade has a single stage, whose score is the detection's output)
    // anything else is re-evaluated with the reference's exact arithmetic
    float lim_reject, lim_skip, lim_noskip, lim_pass;
    float wb[8];                                         // float(w[32] * bias)
    int block_base[SC_PLAN_MAX_SCALES];                  // first stage-0 CTA of every scale inside a frame
    int row_base[SC_PLAN_MAX_SCALES];                    // first lattice row of every scale inside a frame (k_scan_odd)
    float w[SC_F_MAXW][32];
    uint32_t geom[SC_PLAN_MAX_SCALES][SC_F_MAXW][12];    // ScGeom of (this launch's column parity, scale, weak)
};

// Device record of a window that passed stage 0 (or of every window in force_all mode).
//   fs: frame << 8 | scale      yx: gy << 16 | gx (lattice coordinates)
//   rej: stage that rejected it (n_stages = passed all, -1 = still alive)     score: float bits of that stage's score
struct ScRecord {
    uint32_t fs, yx;
    int32_t rej;
    uint32_t score;
};

enum { SC_CNT_VISITED = 0, SC_CNT_PREFILTER = 1, SC_CNT_RAW = 2, SC_CNT_EVALODD = 3, SC_CNT_REACH0 = 4, SC_CNT_STRIDE = 4 + SC_PLAN_MAX_STAGES };

#ifdef __CUDACC__
#define SC_HD __host__ __device__ __forceinline__
#else
#define SC_HD inline
#endif

// layout index (float4 units, low half) of integral pixel (X, Y)
SC_HD long long sc_layout_index(const ScLayout& L, int X, int Y) {
    const int px = X / L.sx, rx = X - px * L.sx, py = Y / L.sy, ry = Y - py * L.sy;
    return (long long)(ry * L.sx + rx) * L.plane4 + (long long)py * L.ppitch + SC_COL(px);
}

#endif
