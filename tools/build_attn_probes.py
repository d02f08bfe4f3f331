"""Build attention-kernel timing-probe variants (libkocr_<name>.so, selected at run time with KOCR_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karanta_ocr_b200 import build
VARIANTS = {
    "p1_noexp": ["KOCR_PROBE=1"],
    "p2_nomax": ["KOCR_PROBE=2"],
    "p3_noexp_nomax": ["KOCR_PROBE=3"],
    "p4_noldtm": ["KOCR_PROBE=4"],
    "p8_nosttm": ["KOCR_PROBE=8"],
    "p15_shell": ["KOCR_PROBE=15"],
    "pp0": ["KOCR_PINGPONG=0"],
    "poly0": ["KOCR_POLY_EVERY=0"],
    "poly2": ["KOCR_POLY_EVERY=2"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(build.build_variant(n, VARIANTS[n]))
