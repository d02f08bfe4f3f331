"""Build attention-kernel timing-probe variants (libkocr_<name>.so, selected at run time with KOCR_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karanta_ocr_b200 import build
VARIANTS = {
    "p1_noexp": ["KOCR_PROBE=1"],
    "p2_nomax": ["KOCR_PROBE=2"],
    "p16_noacc": ["KOCR_PROBE=16"],
    "p32_nosub": ["KOCR_PROBE=32"],
    "p64_nocvt": ["KOCR_PROBE=64"],
    "p114_noacc_nosub_nocvt_nomax": ["KOCR_PROBE=114"],
    "p115_only_ldst": ["KOCR_PROBE=115"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(build.build_variant(n, VARIANTS[n]))
