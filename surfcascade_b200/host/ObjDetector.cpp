// ObjDetector -- command-line detector over the B200 path.
//
// Mirrors the reference's `ObjDetector --detect` branch (ObjDetector.cpp:99-253): model.cfg from the working
// directory (:55), a 40x40 template (:112), every image scanned over the scale ladder, raw windows grouped with
// groupRectangles(wins, weights = 0, scores, 2, 0.2) (:224-225), and the result written as
//     <image name>\n<count>\n<x y w h score>\n...                                  (:228-231)
// The reference hard-codes its Windows paths and the base window 70; here they are arguments:
//     ObjDetector --detect [--model model.cfg] [--base 40] [--out detections.txt] [--batch 16] image.pgm...
// `--train` is the reference's CPU trainer and is not part of this library.
// Images are binary PGM (P5) unless built with -DSC_HAVE_OPENCV, in which case cv::imread is used.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/surfcascade.h"
#include "cvcompat.h"

namespace {

bool read_gray(const std::string& path, cv::Mat* out) {
#ifdef SC_HAVE_OPENCV
    *out = cv::imread(path, cv::IMREAD_GRAYSCALE);
    return !out->empty();
#else
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.good()) return false;
    std::string magic;
    f >> magic;
    if (magic != "P5") return false;
    int vals[3], got = 0;
    while (got < 3 && f.good()) {
        int c = f.peek();
        if (c == '#') { std::string skip; std::getline(f, skip); continue; }
        if (isspace(c)) { f.get(); continue; }
        f >> vals[got++];
    }
    if (got != 3 || vals[2] > 255 || vals[0] < 1 || vals[1] < 1) return false;
    f.get();
    cv::Mat m(vals[1], vals[0], CV_8UC1);
    f.read((char*)m.data, (std::streamsize)vals[0] * vals[1]);
    if (!f.good()) return false;
    *out = m;
    return true;
#endif
}

int usage() {
    std::cout << "Usage: ObjDetector --detect [--model model.cfg] [--base 40] [--out file] [--batch 16] image.pgm..." << std::endl;
    return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 2) return usage();
    if (strcmp(argv[1], "--train") == 0 || strcmp(argv[1], "-t") == 0) {
        std::cerr << "ObjDetector: training is the reference's CPU path and is not part of the B200 library" << std::endl;
        return 2;
    }
    if (strcmp(argv[1], "--detect") != 0 && strcmp(argv[1], "-d") != 0) return usage();

    std::string model = "model.cfg", out_path;
    int base = 40, batch = 16;
    std::vector<std::string> files;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--model" && i + 1 < argc) model = argv[++i];
        else if (a == "--base" && i + 1 < argc) base = atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out_path = argv[++i];
        else if (a == "--batch" && i + 1 < argc) batch = atoi(argv[++i]);
        else files.push_back(a);
    }
    if (files.empty()) return usage();

    sc_handle* h = nullptr;
    if (sc_create(0, &h) != SC_OK) {
        std::cerr << "ObjDetector: no usable CUDA device (there is no CPU path)" << std::endl;
        return 1;
    }
    if (sc_load_model(h, model.c_str(), 40) != SC_OK) {
        std::cerr << "ObjDetector: " << sc_last_error(h) << std::endl;
        return 1;
    }
    std::ofstream of;
    if (!out_path.empty()) of.open(out_path.c_str(), std::ios::binary);
    std::ostream& os = out_path.empty() ? std::cout : of;

    sc_detect_params prm = {};
    prm.base = base; prm.step = 0; prm.scale = 1.1; prm.prefilter = 6; prm.skip_rule = 1; prm.force_all_stages = 0;
    prm.group_threshold = 2; prm.group_eps = 0.2;   // groupRectangles(wins, weights, scores, 2, 0.2), ObjDetector.cpp:224-225, on the device
    std::vector<sc_detection> dets(1 << 16);
    size_t i = 0;
    while (i < files.size()) {
        // batch consecutive images of equal size
        std::vector<cv::Mat> imgs;
        std::vector<const uint8_t*> ptrs;
        size_t j = i;
        for (; j < files.size() && (int)imgs.size() < batch; j++) {
            cv::Mat m;
            if (!read_gray(files[j], &m)) { std::cerr << "ObjDetector: cannot read " << files[j] << std::endl; return 1; }
            if (!imgs.empty() && (m.cols != imgs[0].cols || m.rows != imgs[0].rows || m.step != imgs[0].step)) break;
            imgs.push_back(m);
        }
        for (auto& m : imgs) ptrs.push_back(m.data);
        size_t n = 0;
        int rc = sc_detect(h, ptrs.data(), (int)ptrs.size(), imgs[0].cols, imgs[0].rows, (int)imgs[0].step, &prm, dets.data(), dets.size(), &n, nullptr);
        if (rc == SC_ERR_CAPACITY) {
            dets.resize(n + 16);
            rc = sc_detect(h, ptrs.data(), (int)ptrs.size(), imgs[0].cols, imgs[0].rows, (int)imgs[0].step, &prm, dets.data(), dets.size(), &n, nullptr);
        }
        if (rc != SC_OK) { std::cerr << "ObjDetector: " << sc_last_error(h) << std::endl; return 1; }
        size_t k = 0;
        for (size_t f = 0; f < imgs.size(); f++) {
            std::cout << "Detecting image " << i + f + 1 << '/' << files.size() << std::endl;
            size_t k1 = k;
            while (k1 < n && dets[k1].frame == (int)f) k1++;
            os << files[i + f] << '\n' << (k1 - k) << '\n';   // path, count, then x y w h score (ObjDetector.cpp:228-231)
            for (; k < k1; k++) os << dets[k].x << ' ' << dets[k].y << ' ' << dets[k].l << ' ' << dets[k].l << ' ' << dets[k].score << '\n';
        }
        i += imgs.size();
    }
    sc_destroy(h);
    return 0;
}
