"""Generates tests/golden/model_c1.cfg: the BASELINE config-1 cascade, trained by the REFERENCE's own trainer
(ObjDetector.cpp:66-91 -> CascadeClassifier::Train -> GentleAdaboost::Train -> liblinear) compiled in
oracle/_ref, on the seeded synthetic set of surfcascade_b200/synth.py.  glibc rand() is unseeded in the
reference, hence deterministic per platform.  Run here (needs /root/reference to have built oracle/_ref):

    python tests/golden/train_c1_model.py [n_pos] [n_neg_frames]
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import refbind as R  # noqa: E402
from surfcascade_b200 import synth  # noqa: E402


def main():
    n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    n_neg = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_c1.cfg")
    d = tempfile.mkdtemp(prefix="sc_train_")
    with open(os.path.join(d, "pos.list"), "w") as f:
        for i in range(n_pos):
            R.write_pgm(os.path.join(d, f"p{i:05d}.pgm"), synth.positive(i))
            f.write(f"p{i:05d}.pgm\n")
    with open(os.path.join(d, "neg.list"), "w") as f:
        for i in range(n_neg):
            R.write_pgm(os.path.join(d, f"n{i:05d}.pgm"), synth.negative_frame(i))
            f.write(f"n{i:05d}.pgm\n")
    t = time.time()
    stages = R.train(d, "pos.list", "neg.list", out, verbose=True)
    print(f"trained {stages} stages in {time.time() - t:.1f}s -> {out}")


if __name__ == "__main__":
    main()
