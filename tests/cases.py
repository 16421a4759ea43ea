"""Shared case tables for the parity tests (inputs are seeded and integer-only)."""
import numpy as np

W_DD137, W_CDF53, W_HAAR, W_NONE = 0, 1, 2, 3
C_YCOCG, C_SUBG, C_NONE, C_YCOCG_Q = 0, 1, 2, 3
WR_CLAMP, WR_MIRROR, WR_REPEAT, WR_ZERO = 0, 1, 2, 3

# SURVEY.md Appendix B end-to-end known answers: (w, h, wavelet, q, g, seed, blob_bytes, blob_sha256, decoded_sha256)
KATS = [
    (1024, 1280, W_CDF53, 16, 0, 1, 215699,
     "239ccf2d5fc1dc271f895ced9875638b7dd2bb435bb8a29f94ff8d83a6097bf2",
     "a579ccc1f665fcec969bc9f797e149ff23a72cc46230412bd5bc25a3e5724522"),
    (1632, 2464, W_DD137, 16, 16, 2, 313521,
     "0acb6aa1ceea7a8fa728a4d6d393806159b6ad5c259718783af9293616727d36",
     "dfd18e5fd8e62b7129171094a7e69cef65a01acd4cd1e025404190b6e7d4844b"),
    (1920, 1080, W_DD137, 16, 0, 3, 262567,
     "d1920768f74e3253d5aa4d176fcb0b43dac96583b5df6484bca9cca61edfc16f", None),
    (1920, 1080, W_CDF53, 16, 0, 3, 268750,
     "316e074d81eb34e2ea4fd72416c58e02e1107a7cafafe6dbb249af45b5013249", None),
    (1920, 1080, W_HAAR, 16, 0, 3, 348198,
     "7c7ff0f18ebb954a76103a543343aa242bd85fe56298fb80e5c11ceef58b4dc1", None),
    (1920, 1080, W_CDF53, 0, 0, 3, 3534733,
     "e11f875ceb97680378c169c7289162e80b83734e29b9488b032fc4a6230df527", None),
]
KATS_BIG = [  # GPU-only (seconds to minutes on a CPU)
    (8192, 8192, W_CDF53, 16, 0, 4, 1578381,
     "8abd5bb4edcbbad8c7a3d4194f22e4a108a16d76957c49f6f33bca9fc26be862", None),
    (8192, 8192, W_DD137, 16, 0, 4, 1443551,
     "e22038eb5545ab53709f603e08af26e8a066c343260d0ec5f1bb699c711fcc5f", None),
    (8192, 8192, W_HAAR, 16, 0, 4, 1278745,
     "1e064397a0c8d8d1de6a62cab5e16f11093ae74be5afb789747eea5d123ae925", None),
    (16384, 16384, W_CDF53, 0, 0, 5, 458045756,
     "19a2c7eb8be60101348acec4a2f68105ca48607d72c9caeab6bda5286ba9d930", None),
]


def noise_image(w, h, ch, seed, lo=0, hi=256):
    rs = np.random.RandomState(seed)
    return rs.randint(lo, hi, size=(h, w, ch)).astype(np.uint8)


def smooth_image(w, h, ch, seed):
    """Low-frequency ramps + a little noise: compresses well, exercises long zero runs."""
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    planes = []
    for c in range(ch):
        a, b = rs.randint(1, 5, size=2)
        p = (x * a + y * b + c * 37) // 3 % 256
        p = p + rs.randint(-2, 3, size=(h, w))
        planes.append(np.clip(p, 0, 255))
    return np.stack(planes, axis=-1).astype(np.uint8)


# (w, h, channels) shapes that stress the plus-one rule, DD137->CDF53 fallback (<8), tiny pyramids
SHAPES = [(64, 64, 4), (65, 63, 4), (33, 47, 3), (17, 16, 1), (16, 17, 2), (3, 3, 4), (5, 9, 3), (100, 7, 4),
          (129, 255, 4), (256, 130, 5), (31, 31, 16)]
