"""CPU-only checks of the command-line tools (tools/akoenc.c, tools/akodec.c): they build and link against the
library, accept the reference tools' command lines (tools/options.hpp), and their PNG reader/writer (tools/png_min.c,
standing in for the reference's vendored lodepng) agrees with Pillow and with PNG files the reference's akodec wrote."""
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "_bin")
MODES = {1: "L", 2: "LA", 3: "RGB", 4: "RGBA"}


@pytest.fixture(scope="module")
def tools():
    import ako_b200
    ako_b200.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "tools"), "-s"], check=True)
    return BIN


def run(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True)


def picture(w, h, ch, seed):
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = (x * 3 + y * 2)[..., None] + np.arange(ch) * 40
    return ((base + rs.randint(0, 9, size=(h, w, ch))) % 256).astype(np.uint8)


@pytest.mark.parametrize("shape", [(7, 5, 1), (16, 9, 2), (33, 20, 3), (64, 48, 4), (1, 1, 4), (300, 2, 3)])
def test_png_reader_and_writer_agree_with_pillow(tools, tmp_path, shape):
    w, h, ch = shape
    img = picture(w, h, ch, w + ch)
    src, dst = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    Image.fromarray(img.squeeze(-1) if ch == 1 else img, MODES[ch]).save(src)
    for effort in (1, 7, 10):
        r = run(os.path.join(tools, "png_copy"), src, dst, str(effort))
        assert r.returncode == 0, r.stderr
        assert r.stdout.split() == [str(w), str(h), str(ch)]
        back = np.asarray(Image.open(dst)).reshape(h, w, ch)
        assert np.array_equal(back, img)


def test_png_reader_refuses_what_the_reference_tool_refuses(tools, tmp_path):
    # tools/akoenc.cpp:79-91: palette and 16-bit inputs are errors
    pal = str(tmp_path / "p.png")
    Image.fromarray(picture(8, 8, 3, 1), "RGB").quantize(16).save(pal)
    r = run(os.path.join(tools, "png_copy"), pal, str(tmp_path / "o.png"))
    assert r.returncode != 0 and "Unsupported channels number (3)" in r.stderr
    deep = str(tmp_path / "d.png")
    Image.fromarray(picture(8, 8, 1, 2).squeeze(-1).astype(np.uint16) * 200).save(deep)
    r = run(os.path.join(tools, "png_copy"), deep, str(tmp_path / "o.png"))
    assert r.returncode != 0 and "Unsupported bits per pixel-component (16)" in r.stderr
    junk = tmp_path / "j.png"
    junk.write_bytes(b"not a png at all, just bytes" * 4)
    assert run(os.path.join(tools, "png_copy"), str(junk), str(tmp_path / "o.png")).returncode != 0
    good = str(tmp_path / "g.png")
    Image.fromarray(picture(8, 8, 3, 1), "RGB").save(good)
    data = bytearray(open(good, "rb").read())
    data[45] ^= 0x40  # flip a bit inside IDAT: CRC catches it
    (tmp_path / "c.png").write_bytes(bytes(data))
    assert run(os.path.join(tools, "png_copy"), str(tmp_path / "c.png"), str(tmp_path / "o.png")).returncode != 0


def test_reads_png_written_by_the_reference_decoder(tools, tmp_path, orc):
    """akodec of the unmodified reference (lodepng writer) -> our reader gives the oracle's decoded pixels."""
    import oracle_lib as ol
    ref_dec = os.path.join(ROOT, "oracle", "_ref", "akodec")
    if not os.path.exists(ref_dec):
        pytest.skip("reference tools not built (no /root/reference)")
    img = ol.synth(orc, 97, 61, 3)
    blob, _ = ol.orc_encode(orc, img, wavelet=0, q=8, g=0)
    (tmp_path / "x.ako").write_bytes(blob)
    assert run(ref_dec, "-i", str(tmp_path / "x.ako"), "-o", str(tmp_path / "x.png"), "-quiet").returncode == 0
    r = run(os.path.join(tools, "png_copy"), str(tmp_path / "x.png"), str(tmp_path / "y.png"))
    assert r.returncode == 0, r.stderr
    want, _ = ol.orc_decode(orc, blob)
    # lodepng may have reduced the colour type (opaque RGBA -> RGB): compare in RGBA
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "y.png")).convert("RGBA")), want)


def test_command_lines_of_the_reference_tools(tools):
    enc, dec = os.path.join(tools, "akoenc"), os.path.join(tools, "akodec")
    r = run(enc, "--version")
    assert r.returncode == 0 and r.stdout.startswith("Ako encoding tool v0.2.0\n - libako v0.2.0, format 2")
    r = run(dec, "-v")
    assert r.returncode == 0 and r.stdout.startswith("Ako decoding tool v0.2.0")
    for tool in (enc, dec):
        r = run(tool, "--help")
        assert r.returncode == 0 and r.stdout.startswith("USAGE\n    ako")
        r = run(tool)  # tools/akoenc.cpp:219, akodec.cpp:147
        assert r.returncode == 1 and "No input filename specified" in r.stdout
        r = run(tool, "--frobnicate", "1")
        assert r.returncode == 1 and "Error, unknown option '--frobnicate'." in r.stderr
        r = run(tool, "-i")
        assert r.returncode == 1 and "Error, no value specified for option '-i'." in r.stderr
    for bad in (["-q", "9000"], ["-w", "DCT"], ["-wr", "wobble"], ["-dev-r", "-1"], ["-q", "x"]):
        r = run(enc, *bad, "-i", "nothing.png")
        assert r.returncode == 1 and "Error, invalid value" in r.stderr, bad
    assert run(dec, "-e", "11", "-i", "x.ako").returncode == 1
    r = run(enc, "-i", "/nonexistent/file.png", "-w", "cdf53", "-c", "subtract-g", "-wr", "Repeat", "-d")
    assert r.returncode == 1 and "Png error" in r.stdout  # options parsed (case-insensitively), then the file is missing
    r = run(dec, "-i", "/nonexistent/file.ako")
    assert r.returncode == 1 and "Error at opening file" in r.stdout
