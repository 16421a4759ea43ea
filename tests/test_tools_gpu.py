"""GPU: tools/akoenc + tools/akodec (linked against libako_b200) against the reference's own akoenc / akodec
(oracle/_ref, compiled from /root/reference/tools where they lie) on the same PNG files and command lines:
byte-identical .ako files, identical decoded pixels, identical summary lines."""
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "_bin")
REF = os.path.join(ROOT, "oracle", "_ref")
MODES = {1: "L", 2: "LA", 3: "RGB", 4: "RGBA"}


@pytest.fixture(scope="module")
def tools():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tools"), "-s"], check=True)
    if not os.path.exists(os.path.join(REF, "akoenc")):
        pytest.skip("reference tools not built")
    return BIN


def run(*cmd):
    r = subprocess.run(list(cmd), capture_output=True, text=True)
    assert r.returncode == 0, (cmd, r.stdout, r.stderr)
    return r.stdout


CASES = [
    (640, 360, 4, []),
    (640, 360, 4, ["-q", "0", "-w", "CDF53"]),
    (333, 211, 3, ["-q", "40", "-g", "12", "-w", "HAAR", "-wr", "MIRROR"]),
    (257, 129, 1, ["-c", "NONE", "-wr", "REPEAT", "-chroma-loss", "0"]),
    (200, 300, 2, ["-d", "-q", "8"]),
    (640, 360, 4, ["-dev-r", "20"]),
    (640, 360, 4, ["-dev-r", "1"]),
    (512, 512, 3, ["-dev-r", "8", "-w", "CDF53", "-c", "SUBTRACT-G"]),
    (300, 200, 4, ["-dev-compression", "NONE", "-q", "4"]),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}{''.join(c[3])}")
def test_tools_match_reference_tools(tools, tmp_path, orc, case):
    w, h, ch, flags = case
    img = ol.synth(orc, w, h, 21 + ch)[:, :, :ch] if ch != 2 else ol.synth(orc, w, h, 23)[:, :, 2:]
    img = np.ascontiguousarray(img)
    png = str(tmp_path / "in.png")
    Image.fromarray(img.squeeze(-1) if ch == 1 else img, MODES[ch]).save(png)
    ours, theirs = str(tmp_path / "ours.ako"), str(tmp_path / "ref.ako")
    out_ours = run(os.path.join(tools, "akoenc"), "-i", png, "-o", ours, "-ch", *flags)
    out_ref = run(os.path.join(REF, "akoenc"), "-i", png, "-o", theirs, "-ch", *flags)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    assert out_ours.strip().splitlines()[-1] == out_ref.strip().splitlines()[-1]  # "(adler) x kB -> y kB, ratio, bpp"

    p_ours, p_ref = str(tmp_path / "ours.png"), str(tmp_path / "ref.png")
    out_ours = run(os.path.join(tools, "akodec"), "-i", theirs, "-o", p_ours, "-ch")
    out_ref = run(os.path.join(REF, "akodec"), "-i", theirs, "-o", p_ref, "-ch")
    assert out_ours.strip().splitlines()[-1] == out_ref.strip().splitlines()[-1]
    a = np.asarray(Image.open(p_ours))
    assert a.reshape(h, w, ch).shape == (h, w, ch)
    assert np.array_equal(np.asarray(Image.open(p_ours).convert("RGBA")), np.asarray(Image.open(p_ref).convert("RGBA")))


def test_tools_benchmark_output(tools, tmp_path, orc):
    """-b: stage stopwatches driven by the library's events, in the reference's order (tools/benchmark.hpp:73-90)."""
    png = str(tmp_path / "in.png")
    Image.fromarray(ol.synth(orc, 320, 240, 5), "RGBA").save(png)
    ako = str(tmp_path / "x.ako")
    lines = run(os.path.join(tools, "akoenc"), "-i", png, "-o", ako, "-b").splitlines()
    assert lines[0].startswith("Benchmark:")
    assert [l.split(":")[0] for l in lines[1:5]] == [" - Format", " - Wavelet transformation", " - Compression", " - Total"]
    lines = run(os.path.join(tools, "akodec"), "-i", ako, "-b").splitlines()
    assert [l.split(":")[0] for l in lines[1:5]] == [" - Compression", " - Wavelet transformation", " - Format", " - Total"]
    assert run(os.path.join(tools, "akoenc"), "-i", png, "-quiet") == ""
