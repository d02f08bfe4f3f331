"""Static checks on the built library's machine code (no GPU needed): the hot kernels really are tcgen05 / TMA code, and
the attention kernel's exponent phase sits between the two named-barrier operations that hand the MUFU unit from one
query tile's softmax warp to the other's. ptxas orders bar.sync / bar.arrive only against memory operations; the source
pins them with data dependencies (kocr_attention.cu), and this test is what notices if a compiler change undoes that."""
import os
import re
import shutil
import subprocess

import pytest

from karanta_ocr_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or libkocr.so not available")
    out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].strip())
    return funcs


def _of(funcs, fragment):
    hits = [v for k, v in funcs.items() if fragment in k]
    assert hits, f"no kernel matching {fragment}"
    return hits


def test_only_sm100a_code_is_embedded():
    out = subprocess.run([CUOBJDUMP, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_gemm_and_attention_use_tcgen05_tmem_tma(sass):
    for frag in ("gemm_kernel", "attention_kernel", "attention3_kernel"):
        for body in _of(sass, frag):
            text = "\n".join(body)
            assert "UTCHMMA" in text, f"{frag}: no tcgen05.mma"
            assert "UTMALDG" in text, f"{frag}: no TMA tensor load"
            assert "LDTM" in text, f"{frag}: no tcgen05.ld"
            assert "HMMA." not in text, f"{frag}: legacy mma.sync found"


def test_preprocess_uses_bulk_store(sass):
    for body in _of(sass, "preprocess_kernel"):
        assert any(op.startswith("UBLKCP") or "UBLKCP" in op for op in body), "no cp.async.bulk store"


def test_attention_exponent_phase_is_between_the_hand_over_barriers(sass):
    for body in _of(sass, "attention_kernelILb0"):  # the two-tile (full attention) shape; windowed CTAs hold one tile
        sync = [i for i, op in enumerate(body) if re.match(r"BAR\.SYNC\S* R\d+, 0x40", op)]
        arv = [i for i, op in enumerate(body) if re.match(r"(@!?P\d+ )?BAR\.ARV R\d+, 0x40", op)]
        assert len(sync) == 1 and len(arv) == 2, (sync, arv)  # wait for the token; initial hand-over + per-sub-step hand-over
        mufu = [i for i, op in enumerate(body) if "MUFU.EX2" in op]
        inside = [i for i in mufu if sync[0] < i < arv[1]]
        after = [i for i in mufu if i > arv[1]]
        before = [i for i in mufu if i < sync[0]]
        # 60 MUFU exponentials per 80-key sub-step (20 more go through the FMA-pipe polynomial) + 1 for the rescale factor
        assert len(mufu) == 61, len(mufu)
        assert len(before) <= 1, "exponentials were hoisted above the token wait"
        assert len(inside) >= 45 and len(after) <= 15, (len(inside), len(after))
