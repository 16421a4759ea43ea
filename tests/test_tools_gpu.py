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


LIB = os.path.join(ROOT, "ako_b200", "libako_b200.so")


@pytest.mark.parametrize("flags", [[], ["-q", "0", "-w", "CDF53"], ["-q", "30", "-g", "10", "-w", "HAAR"], ["-dev-r", "12"],
                                   ["-wr", "MIRROR", "-c", "SUBTRACT-G", "-q", "8"]], ids=lambda f: "".join(f) or "defaults")
def test_reference_tools_preloaded_with_libako_b200(tools, tmp_path, orc, flags):
    """The drop-in claim of INTEGRATION.md section 1 on the reference's OWN callers: its akoenc / akodec binaries
    (tools/akoenc.cpp:118-213, tools/akodec.cpp:139), unmodified, with LD_PRELOAD=libako_b200.so so that every
    library/ako.h symbol resolves to the CUDA implementation: same .ako bytes, same pixels, same summary lines as
    the same binaries on their own library; -b shows the stages timed through the events callback."""
    img = ol.synth(orc, 520, 384, 31)
    png = str(tmp_path / "in.png")
    Image.fromarray(img, "RGBA").save(png)
    env = dict(os.environ, LD_PRELOAD=LIB)

    def run_env(env_, *cmd):
        r = subprocess.run(list(cmd), capture_output=True, text=True, env=env_)
        assert r.returncode == 0, (cmd, r.stdout, r.stderr)
        return r.stdout

    a_ref, a_pre = str(tmp_path / "ref.ako"), str(tmp_path / "pre.ako")
    out_ref = run_env(os.environ, os.path.join(REF, "akoenc"), "-i", png, "-o", a_ref, "-ch", *flags)
    out_pre = run_env(env, os.path.join(REF, "akoenc"), "-i", png, "-o", a_pre, "-ch", *flags)
    assert open(a_ref, "rb").read() == open(a_pre, "rb").read()
    assert out_ref.strip().splitlines()[-1] == out_pre.strip().splitlines()[-1]
    p_ref, p_pre = str(tmp_path / "ref.png"), str(tmp_path / "pre.png")
    out_ref = run_env(os.environ, os.path.join(REF, "akodec"), "-i", a_ref, "-o", p_ref, "-ch")
    out_pre = run_env(env, os.path.join(REF, "akodec"), "-i", a_ref, "-o", p_pre, "-ch")
    assert out_ref.strip().splitlines()[-1] == out_pre.strip().splitlines()[-1]
    assert np.array_equal(np.asarray(Image.open(p_ref)), np.asarray(Image.open(p_pre)))
    # the preloaded library really is the one at work: its events drive the tool's stopwatches, and the loader maps it
    bench = run_env(env, os.path.join(REF, "akoenc"), "-i", png, "-o", a_pre, "-b", *flags)
    assert "Benchmark" in bench
    maps = subprocess.run(["bash", "-c", f"LD_PRELOAD={LIB} LD_DEBUG=libs {os.path.join(REF, 'akodec')} -i {a_ref} -o {p_pre} 2>&1 | grep -c libako_b200"],
                          capture_output=True, text=True)
    assert int(maps.stdout.strip() or 0) > 0
