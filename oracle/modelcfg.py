"""TEST INFRASTRUCTURE ONLY -- the oracle's reading of a reference model.cfg.

Restates what the detector takes from `Model::Load` (Model.cpp:97-193): libconfig 1.4.9 text
(grammar: libconfig/grammar.c, scanner.c -- groups {}, lists (), arrays [], `name = value;` or
`name : value;`), floats parsed as C doubles (libconfig/scanner.c:1146 `atof`) and cast to float32
on assignment to `LogisticRegression::w` / `theta` (Model.cpp:148,172-175; libconfigcpp.c++:710-716).
Pinned against the reference's own loader through oracle/_ref `ref_model_load` in tests/test_oracle_vs_ref.py.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

_TOKEN = re.compile(r"""
    (?P<ws>\s+|\#[^\n]*|//[^\n]*|/\*.*?\*/)
  | (?P<float>[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+))
  | (?P<hex>0[xX][0-9a-fA-F]+L{0,2})
  | (?P<int>[-+]?\d+L{0,2})
  | (?P<bool>[Tt][Rr][Uu][Ee]|[Ff][Aa][Ll][Ss][Ee])
  | (?P<name>[A-Za-z\*][-A-Za-z0-9_\*]*)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<punct>[{}()\[\]=:;,])
""", re.VERBOSE | re.DOTALL)


def _tokens(text: str):
    pos = 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"model.cfg: cannot tokenise at offset {pos}: {text[pos:pos + 20]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            yield kind, m.group(kind)
    yield "eof", ""


class _Parser:
    def __init__(self, text: str):
        self.toks = list(_tokens(text))
        self.i = 0

    def peek(self):
        return self.toks[self.i]

    def take(self, kind=None, val=None):
        k, v = self.toks[self.i]
        if (kind and k != kind) or (val and v != val):
            raise ValueError(f"model.cfg: expected {kind or val}, got {k} {v!r}")
        self.i += 1
        return v

    def settings(self, end):
        out = {}
        while not (self.peek()[0] == end[0] and (end[1] is None or self.peek()[1] == end[1])):
            name = self.take("name")
            if self.peek()[1] not in ("=", ":"):
                raise ValueError("model.cfg: expected = or :")
            self.take("punct")
            out[name] = self.value()
            if self.peek() in (("punct", ";"), ("punct", ",")):
                self.take("punct")
        return out

    def value(self):
        k, v = self.peek()
        if k == "punct" and v == "{":
            self.take()
            g = self.settings(("punct", "}"))
            self.take("punct", "}")
            return g
        if k == "punct" and v in "([":
            close = ")" if v == "(" else "]"
            self.take()
            items = []
            while self.peek() != ("punct", close):
                items.append(self.value())
                if self.peek() == ("punct", ","):
                    self.take()
            self.take("punct", close)
            return items
        self.take()
        if k == "float":
            return float(v)
        if k == "int":
            return int(v.rstrip("L"))
        if k == "hex":
            return int(v.rstrip("L"), 16)
        if k == "bool":
            return v.lower() == "true"
        if k == "str":
            return v[1:-1]
        raise ValueError(f"model.cfg: unexpected token {k} {v!r}")


def parse(text: str) -> dict:
    p = _Parser(text)
    root = p.settings(("eof", None))
    return root


@dataclass
class Cascade:
    """Flattened cascade, the arrays `so_cascade` / `sc_cascade_desc` point at."""
    theta: np.ndarray        # f32 [n_stages]
    n_weak: np.ndarray       # i32 [n_stages]
    patch_index: np.ndarray  # i32 [total]
    w: np.ndarray            # f32 [total][33]
    bias: np.ndarray         # f64 [total]

    @property
    def n_stages(self) -> int:
        return len(self.theta)


def load(path: str) -> Cascade:
    """Model::Load semantics: a missing setting silently ends the load (Model.cpp:188-191), so stages /
    weak classifiers parsed before the gap are kept; a float field holding an int raises, like the
    uncaught SettingTypeException would abort the reference (libconfigcpp.c++:1137-1145)."""
    with open(path, "r") as f:
        root = parse(f.read())
    theta, n_weak, pidx, w, bias = [], [], [], [], []

    def need_float(v):
        if not isinstance(v, float):
            raise TypeError("model.cfg: float setting holds a non-float (reference would abort)")
        return v

    try:
        cc = root["cascade_classifier"]
        for key in ("max_stages_num", "FPR_target", "TPR_min_perstage", "FPR", "TPR"):
            cc[key]
        for st in cc["stage_classifiers"]:
            for key in ("search_step", "auc_step", "TPR_min", "n_total", "n_pos", "n_neg", "FPR", "TPR"):
                st[key]
            th = np.float32(need_float(st["theta"]))
            for key in ("total_AUC_score", "sample_num", "max_iters"):
                st[key]
            # the stage object exists before its weak classifiers are read, but it is only pushed to the
            # cascade after all of them parsed (Model.cpp:189); a gap inside drops the whole stage
            sw, sp, sb = [], [], []
            for wk in st["weak_classifiers"]:
                pi = wk["patch_index"]
                for key in ("eps", "C", "nr_class", "nr_feature"):
                    wk[key]
                b = need_float(wk["bias"])
                ws = [np.float32(need_float(x)) for x in wk["w"]]
                wk["label"][1]
                sw.append(ws); sp.append(pi); sb.append(b)
            theta.append(th); n_weak.append(len(sw)); w += sw; pidx += sp; bias += sb
    except (KeyError, IndexError):
        pass
    return Cascade(np.array(theta, np.float32), np.array(n_weak, np.int32), np.array(pidx, np.int32),
                   np.array(w, np.float32).reshape(-1, 33), np.array(bias, np.float64))
