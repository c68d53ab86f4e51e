// Host-side geometry and model flattening shared by the C-ABI and the kept C++ classes.
// Each function restates, in this repo's own code, the reference arithmetic it names (paths relative to
// /root/reference/ObjDetector); float32 operations are kept float32 so that truncations land identically.
#ifndef SC_HOST_H
#define SC_HOST_H

#include <string>
#include <vector>

#include "../../include/surfcascade.h"
#include "../csrc/sc_plan.h"

namespace sc_host {

// ExtractPatches, FeatureExtractors/DenseSURFFeatureExtractor.cpp:49-63
void pool_patches(int tw, int th, std::vector<sc_rect>* out);

// ProjectPatches at window origin (0,0), DenseSURFFeatureExtractor.cpp:486-508
sc_rect project_patch(int tmpl, int l, const sc_rect& patch);

// ProjectPatches + GetRectsFromPatch (:360-377) folded into the layout offsets of the corner lattice (relative to a
// window origin on the layout's lattice); false if the projected patch is not 2x2 / 4x1 / 1x4 cells
// x0 = (window origin x) mod L.sx: the origin's column residue (0 for even lattice columns and for the hooks,
// `step` for odd lattice columns of a detection plan)
bool project_geom(int tmpl, int l, const sc_rect& patch, const ScLayout& L, int x0, ScGeom* g);

// Integral-image layout deinterleaved by sx columns and sy rows (sc_plan.h)
ScLayout make_layout(int W, int H, int sx, int sy, int min_hp = 256);

// Window sides of the scale loop, ObjDetector.cpp:174,180
void scale_ladder(int W, int H, int base, double scale, std::vector<int>* sides);

// cv::groupRectangles(rects, weights = 0.., scores, thr, eps) as called at ObjDetector.cpp:224-225 (in place)
void group_rectangles(std::vector<sc_rect>* rects, std::vector<double>* scores, int thr, double eps);

struct FlatCascade {
    std::vector<float> theta;
    std::vector<int> n_weak;
    std::vector<sc_rect> rects;
    std::vector<int> patch_index;
    std::vector<float> w;       // [total][33]
    std::vector<double> bias;
};

// Model::Load then Model::Save (round trip through this library's reader and writer)
bool resave_model(const std::string& in_cfg, const std::string& out_cfg);

// Model::Load + GetFittedPatchIndexes + dense_patches[patch_index], ObjDetector.cpp:108-130
bool load_flat_cascade(const std::string& model_cfg, int tmpl, FlatCascade* out, std::string* why);

}  // namespace sc_host

#endif
