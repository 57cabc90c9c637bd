"""Host-side builders of shared-memory operand images / descriptors for rz_umma_probe."""
import numpy as np


def sw128(row, byte_in_row):
    return row * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15))


def image_k_major(x: np.ndarray) -> np.ndarray:
    """x [R, K] fp16, K multiple of 64 -> bytes: chunk c = [R x 128B] swizzled."""
    R, K = x.shape
    assert K % 64 == 0
    img = np.zeros(R * K * 2, dtype=np.uint8)
    raw = x.astype(np.float16).view(np.uint8).reshape(R, K * 2)
    for c in range(K // 64):
        base = c * R * 128
        for r in range(R):
            for u in range(8):  # 16-byte units
                src = raw[r, c * 128 + u * 16: c * 128 + u * 16 + 16]
                o = base + sw128(r, u * 16)
                img[o:o + 16] = src
    return img


def image_mn_major(x: np.ndarray) -> np.ndarray:
    """x [MN, K] fp16 (logical), stored MN-contiguous: block j = [K rows x 128B] swizzled."""
    MN, K = x.shape
    assert MN % 64 == 0
    img = np.zeros(MN * K * 2, dtype=np.uint8)
    xt = np.ascontiguousarray(x.astype(np.float16).T)  # [K, MN]
    raw = xt.view(np.uint8).reshape(K, MN * 2)
    for j in range(MN // 64):
        base = j * K * 128
        for k in range(K):
            for u in range(8):
                src = raw[k, j * 128 + u * 16: j * 128 + u * 16 + 16]
                o = base + sw128(k, u * 16)
                img[o:o + 16] = src
    return img


def smem_desc(start_bytes=0, lbo=0, sbo=1024, layout=2, version=1):
    d = (start_bytes >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= version << 46
    d |= layout << 61
    return d


def idesc_f16(m, n, a_mn=0, b_mn=0):
    return (1 << 4) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def lane_of_row(m_rows: int, r: int) -> int:
    """TMEM lane holding accumulator row r (cta_group::1)."""
    if m_rows == 128:
        return r
    return (r % 16) + 32 * (r // 16)  # M = 64: 16 lanes per 32-lane sub-partition
