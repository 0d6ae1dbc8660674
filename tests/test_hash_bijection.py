"""The hub tier's multi-pass selection (lp_sweep.cuh: sweep_hub_select) terminates because lowbias32 is a BIJECTION
on 32-bit words: two distinct labels differ in at least one hash bit, so splitting a bucket into 2, 4, ... hash
classes eventually separates any set of labels. The commit priorities (bijective32) rely on the same property
(unique priorities). Both are xorshift / odd-multiply chains; this test inverts them step by step on random and
edge-case words (numpy restatement of kaminpar_b200/csrc/lp_device.cuh:22-45, test infrastructure only)."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mul(x, c):
    return (x.astype(np.uint64) * np.uint64(c) & M32).astype(np.uint32)


def lowbias32(x):
    x = x ^ (x >> np.uint32(16))
    x = mul(x, 0x7FEB352D)
    x = x ^ (x >> np.uint32(15))
    x = mul(x, 0x846CA68B)
    return x ^ (x >> np.uint32(16))


def bijective32(x, base):
    x = x ^ np.uint32(base)
    x = mul(x, 0x9E3779B1)
    x = x ^ (x >> np.uint32(15))
    x = mul(x, 0x85EBCA77)
    x = x ^ (x >> np.uint32(13))
    x = mul(x, 0xC2B2AE3D)
    return x ^ (x >> np.uint32(16))


def inv_xorshift(y, s):
    x = y.copy()
    for _ in range(32 // s + 1):  # x = y ^ (x >> s) converges from the top bits down
        x = y ^ (x >> np.uint32(s))
    return x


def inv_mul(c):
    assert c & 1
    return pow(c, -1, 1 << 32)


def words():
    rng = np.random.default_rng(11)
    edge = np.array([0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, 0x00FFFFFF, 0x01000000], np.uint32)
    return np.concatenate([edge, rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32),
                           np.arange(1 << 16, dtype=np.uint32)])


def test_lowbias32_is_invertible():
    x = words()
    y = lowbias32(x)
    z = inv_xorshift(y, 16)
    z = mul(z, inv_mul(0x846CA68B))
    z = inv_xorshift(z, 15)
    z = mul(z, inv_mul(0x7FEB352D))
    z = inv_xorshift(z, 16)
    assert np.array_equal(z, x)
    assert len(np.unique(y)) == len(np.unique(x))


def test_bijective32_is_invertible():
    x = words()
    for base in (0, 0xDEADBEEF):
        y = bijective32(x, base)
        z = inv_xorshift(y, 16)
        z = mul(z, inv_mul(0xC2B2AE3D))
        z = inv_xorshift(z, 13)
        z = mul(z, inv_mul(0x85EBCA77))
        z = inv_xorshift(z, 15)
        z = mul(z, inv_mul(0x9E3779B1))
        z = z ^ np.uint32(base)
        assert np.array_equal(z, x)
