"""Seeded synthetic raw images for the preprocessing parity tests (chest-X-ray-like smooth field + noise,
in the integer / float types DICOM and PNG sources produce).  Shared by the golden generator and the tests."""
import numpy as np

CASES = {
    # name: raw shape, dtype, value range, processor size
    "cxr1024_u8": dict(shape=(1024, 1024), dtype="uint8", hi=255, size=(518, 518), seed=1),      # config C1
    "dicom_u16": dict(shape=(700, 900), dtype="uint16", hi=4095, size=(518, 518), seed=2),
    "small_u8_upsample": dict(shape=(300, 420), dtype="uint8", hi=255, size=(518, 518), seed=3),
    "rgb_u8": dict(shape=(333, 777, 3), dtype="uint8", hi=255, size=(224, 224), seed=4),
    "padchest_i32": dict(shape=(2021, 2500), dtype="int32", hi=65535, size=(518, 518), seed=5),
    "float_identity_w": dict(shape=(600, 518), dtype="float32", hi=None, size=(518, 518), seed=6),
    "signed_i16": dict(shape=(512, 640), dtype="int16", hi=3000, size=(518, 518), seed=7),
    "rgb_u8_mild": dict(shape=(600, 700, 3), dtype="uint8", hi=255, size=(518, 518), seed=10),   # short windows, 3 channels
    "u16_tall": dict(shape=(1100, 640), dtype="uint16", hi=65535, size=(518, 320), seed=11),      # odd band edges
    "odd_sizes_u8": dict(shape=(333, 257), dtype="uint8", hi=255, size=(199, 131), seed=12),      # partial column blocks / bands
    "constant": dict(shape=(64, 64), dtype="uint8", hi=0, size=(37, 41), seed=8),
    "levels": dict(shape=(16, 16), dtype="uint8", hi=-1, size=(16, 16), seed=9, keep_pv=True),
}


def make_raw(spec) -> np.ndarray:
    rng = np.random.default_rng(spec["seed"])
    shape, dt = spec["shape"], np.dtype(spec["dtype"])
    if spec["hi"] == -1:                       # every uint8 level once: exercises the whole normalise table
        return np.arange(256, dtype=np.uint8).reshape(16, 16)
    if spec["hi"] == 0:
        return np.full(shape, 7, dtype=dt)
    if spec["hi"] is None:
        return (rng.standard_normal(shape) * 500.0).astype(dt)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    base = np.sin(yy / 97.0) * np.cos(xx / 61.0) * 0.5 + 0.5
    if len(shape) == 3:
        base = base[..., None] * np.array([1.0, 0.8, 0.6])
    hi = spec["hi"]
    raw = np.clip(base * hi * 0.9 + rng.integers(0, hi // 10 + 1, size=shape), 0, hi)
    if dt == np.int16:
        raw = raw - hi // 2
    return raw.astype(dt)
