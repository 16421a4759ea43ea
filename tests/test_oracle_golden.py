"""Oracle vs the committed golden vectors (generated from the reference by tests/golden/make_golden.py).
Runs without /root/reference and without a GPU."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden.make_golden import make_input

with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as f:
    GOLDEN = json.load(f)


def _id(c):
    return f"{c['kind']}-{c['w']}x{c['h']}x{c['ch']}-w{c['wavelet']}-r{c['wrap']}-q{c['q']}g{c['g']}-t{c['tiles']}-c{c['color']}{c['discard']}"


@pytest.mark.parametrize("c", GOLDEN, ids=_id)
def test_oracle_reproduces_golden(orc, c):
    img = make_input(orc, c)
    assert hashlib.sha256(img.tobytes()).hexdigest() == c["input_sha256"]
    blob, st = ol.orc_encode(orc, img, wavelet=c["wavelet"], wrap=c["wrap"], q=c["q"], g=c["g"], tiles=c["tiles"],
                             color=c["color"], discard=c["discard"], chroma_loss=c["chroma_loss"])
    assert st == 0 and len(blob) == c["blob_len"]
    assert hashlib.sha256(blob).hexdigest() == c["blob_sha256"]
    if "blob_b64" in c:
        assert blob == base64.b64decode(c["blob_b64"])
    dec, st = ol.orc_decode(orc, blob)
    assert st == 0 and hashlib.sha256(dec.tobytes()).hexdigest() == c["decoded_sha256"]
    if c["q"] == 0 and c["g"] == 0 and c["discard"] == 0:
        assert np.array_equal(dec, img)
