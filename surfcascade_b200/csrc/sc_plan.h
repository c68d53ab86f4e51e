// Host/device-shared plan structures for the SURF-cascade scan (sm_100a).
#ifndef SC_PLAN_H
#define SC_PLAN_H

#include <stdint.h>

#define SC_PLAN_MAX_SCALES 64
#define SC_PLAN_MAX_STAGES 16
#define SC_W_PITCH 36          // floats per weak classifier in device memory: w[0..32] + 3 pad (16 B aligned rows)

// Stage-0 tile: 64 x 16 window origins per CTA (two 32-wide bitmask words per tile row).
#define SC_TILE_X 64
#define SC_TILE_Y 16
#define SC_TILE_THREADS 256

// Integral strips: one warp walks one 32-column strip down the frame.
#define SC_STRIP 32

struct ScScale {
    int l;            // window side (int)(base * scale^i), ObjDetector.cpp:180
    int nx, ny;       // window origins per row / rows on the step lattice
    int wpr;          // 32-bit bitmask words per lattice row = ceil(nx / 32)
    float thr;        // float(l * l * prefilter), ObjDetector.cpp:188
    int tiles_x;      // stage-0 tiles per row of tiles
    int block_base;   // first stage-0 CTA of this scale inside a frame
    int word_base;    // first bitmask word of this scale inside a frame
    int row_base;     // first lattice row of this scale inside a frame (replay threads)
    int pad;
};

struct ScPlan {
    int W, H, pitch;              // pitch = W + 1 integral pixels per row
    int step;
    int n_scales;
    int n_stages;
    int total_weak;
    int use_prefilter, skip_rule, force_all;
    int blocks_per_frame;         // stage-0 CTAs per frame
    int words_per_frame;          // bitmask words per frame
    int rows_per_frame;           // lattice rows per frame (all scales)
    int n_strips;                 // ceil(W / 32)
    long long windows_per_frame;  // grid windows per frame
    long long frame_stride4;      // float4 elements between consecutive frames' integrals
    float theta[SC_PLAN_MAX_STAGES];
    int n_weak[SC_PLAN_MAX_STAGES];
    int weak_base[SC_PLAN_MAX_STAGES];
    // multi == 2 for a window rejected at stage p with score s  <=>  ((double)s + p + 1) / n_stages < 0.5
    ScScale sc[SC_PLAN_MAX_SCALES];
};

// Per (scale, weak classifier) projected geometry, relative to the window origin, in integral pixels.
//   off    = oy * pitch + ox           first corner
//   along  = pixel stride between consecutive corners along the cell chain (square: ce; wide: ce; tall: ce * pitch)
//   across = pixel stride to the second line of corners                    (square: ce * pitch; wide: ce * pitch; tall: ce)
//   shape  = 0 square 2x2 (3 x 3 corners), 1 long 4x1 / 1x4 (2 x 5 corners)
struct ScGeom {
    int off, along, across, shape;
};

// Device record of a window that passed stage 0 (or of every window in force_all mode).
//   x: frame << 8 | scale      y: gy << 16 | gx (lattice coordinates)
//   z: stage that rejected it (n_stages = passed all, -1 = still alive)     w: float bits of the score at that stage
struct ScRecord {
    uint32_t fs, yx;
    int32_t rej;
    uint32_t score;
};

enum { SC_CNT_VISITED = 0, SC_CNT_PREFILTER = 1, SC_CNT_RAW = 2, SC_CNT_REACH0 = 3, SC_CNT_STRIDE = 3 + SC_PLAN_MAX_STAGES };

#endif
