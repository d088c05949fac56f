"""Host-side operand packing of the all-tensor-core EPI kernel (lfsr_mel_epi_pack): runs without a GPU.
The image holds, pre-swizzled for the kernel's UMMA descriptors, one [32 out][16 in] fp16 block per depthwise tap
(dw[tap][c] * pw_branch[c][o], K-major SWIZZLE_32B), the extra-channel block, the fuse matrix (SWIZZLE_128B) and the fp32 tap
weights of channels 16, 17 (MyEfficientLFNet.py:278-327 is the block being packed)."""
import numpy as np
import torch

import lfsr_b200
from lfsr_b200 import _native as N

EC = 18


def _f16(v):
    return np.float16(np.clip(np.float32(v), -65504.0, 65504.0))


def test_mel_epi_pack_layout_and_rounding():
    lib = N.load()
    klen, ntap = 11, 31
    assert lib.lfsr_mel_epi_pack_bytes(klen) == ntap * 1024 + 3072 + 4096 + 256
    assert lib.lfsr_mel_epi_pack_bytes(10) == 0 and lib.lfsr_mel_epi_pack_bytes(17) == 0     # even / too many taps
    torch.manual_seed(0)
    w = (torch.rand(ntap * EC + 6 * EC * EC) - 0.5) * 0.6
    w[5], w[7], w[9] = 1e-6, 70000.0, 3e-8              # a subnormal product, an overflow (saturates), an underflow
    img = torch.zeros(lib.lfsr_mel_epi_pack_bytes(klen), dtype=torch.uint8)
    assert lib.lfsr_mel_epi_pack(w.data_ptr(), img.data_ptr(), klen) == 0
    a = img.numpy()
    dw = w[:ntap * EC].view(ntap, EC)
    pw = w[ntap * EC:ntap * EC + 3 * EC * EC].view(3, EC, EC)
    fu = w[ntap * EC + 3 * EC * EC:].view(3 * EC, EC)
    half_at = lambda blk, off: np.frombuffer(blk[off:off + 2].tobytes(), dtype=np.float16)[0]
    for t in range(ntap):
        br = 0 if t < klen else (1 if t < 2 * klen else 2)
        blk = a[t * 1024:(t + 1) * 1024]
        for n in range(32):
            for k in range(16):
                row = n * 32
                off = row + (((k >> 3) ^ ((row >> 7) & 1)) << 4) + (k & 7) * 2
                want = _f16((dw[t, k] * pw[br, k, n]).item()) if n < EC else np.float16(0)
                assert half_at(blk, off) == want, (t, n, k)
    bex = a[ntap * 1024:ntap * 1024 + 3072]
    for br in range(3):
        for o in range(EC):
            for j in range(2):
                n, k = 32 * br + o, 2 * br + j
                row = n * 32
                off = row + (((k >> 3) ^ ((row >> 7) & 1)) << 4) + (k & 7) * 2
                assert half_at(bex, off) == _f16(pw[br, 16 + j, o].item())
    b2 = a[ntap * 1024 + 3072:ntap * 1024 + 3072 + 4096]
    for n in range(32):
        for k in range(64):
            off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2
            want = _f16(fu[k, n].item()) if (n < EC and k < 3 * EC) else np.float16(0)
            assert half_at(b2, off) == want, (n, k)
    exw = np.frombuffer(a[ntap * 1024 + 3072 + 4096:ntap * 1024 + 3072 + 4096 + ntap * 8].tobytes(), dtype=np.float32).reshape(ntap, 2)
    assert np.array_equal(exw, dw[:, 16:18].numpy())


def test_mel_epi_pack_rejects_bad_arguments():
    lib = N.load()
    img = torch.zeros(64, dtype=torch.uint8)
    assert lib.lfsr_mel_epi_pack(None, img.data_ptr(), 11) != 0
    assert lib.lfsr_mel_epi_pack(img.data_ptr(), img.data_ptr(), 10) != 0
    assert b"lfsr_mel_epi_pack" in lib.lfsr_last_error()
