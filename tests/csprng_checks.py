"""CSPRNG parity cases shared by the host-emulation (CPU) and the GPU tests: the C ABI against
oracle/csprng.py on the same states.  `lib` is a _native.Lib, buffers are numpy arrays (emulation) or
torch CUDA tensors."""

import numpy as np

from oracle import csprng as oc


class Buf:
    def __init__(self, use_torch, device=0):
        self.use_torch, self.device = use_torch, device
        if use_torch:
            import torch

            self.torch = torch

    def dev(self, a, dtype=np.int64):
        a = np.ascontiguousarray(a, dtype=dtype)
        return self.torch.from_numpy(a).to(f"cuda:{self.device}") if self.use_torch else a.copy()

    def empty(self, *shape):
        if self.use_torch:
            return self.torch.empty(shape, dtype=self.torch.int64, device=f"cuda:{self.device}")
        return np.empty(shape, dtype=np.int64)

    def ptr(self, t):
        return t.data_ptr() if self.use_torch else t.ctypes.data

    def host(self, t):
        return t.cpu().numpy() if self.use_torch else np.asarray(t)


def make_states(rows, key, nonce, first_counter=0, high=0):
    st = np.zeros((rows, 16), dtype=np.int64)
    st[:, 0:4] = oc.SIGMA_WORDS
    st[:, 4:12] = key
    st[:, 12] = (first_counter + np.arange(rows)) & 0xFFFFFFFF
    st[:, 13] = high
    st[:, 14:] = nonce
    return st


KEY = [0x03020100, 0x07060504, 0x0B0A0908, 0x0F0E0D0C, 0x13121110, 0x17161514, 0x1B1A1918, 0x1F1E1D1C]
NONCE = [0x4A000000, 0x00000000]


def check_chacha20(lib, b: Buf):
    # RFC 8439 2.3.2 known answer through the kernel: counter word 12 = 1, word 13 = 0x09000000
    st = make_states(1, KEY, [0x4A000000, 0])
    st[0, 12], st[0, 13] = 1, 0x09000000
    d, out = b.dev(st), b.empty(1, 16)
    lib.check(lib.tb200_chacha20(b.device, b.ptr(d), 1, b.ptr(out), 5, None), "chacha20")
    want = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
            0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert list(b.host(out)[0]) == want, "RFC 8439 block function known answer"
    assert list(b.host(d)[0, 12:14]) == [6, 0x09000000]
    # many rows, counters that wrap 2^32 when stepped (carry into word 13), two consecutive calls
    rows = 777
    st = make_states(rows, [7, 8, 9, 10, 11, 12, 13, 0xFFFFFFFF], [5, 6], first_counter=0xFFFFFF00)
    ref = st.copy()
    d, out = b.dev(st), b.empty(rows, 16)
    for call in range(2):
        lib.check(lib.tb200_chacha20(b.device, b.ptr(d), rows, b.ptr(out), 1000, None), "chacha20")
        want = oc.chacha20(ref, 1000)
        assert np.array_equal(b.host(out), want), f"chacha20 blocks, call {call}"
        assert np.array_equal(b.host(d), ref), f"chacha20 stepped states, call {call}"
    assert ref[:, 13].max() == 1  # the carry path ran


def check_randint(lib, b: Buf):
    q = np.array([3, 2, 1000003, (1 << 40) - 87, (1 << 60) - 93, (1 << 64) - 59], dtype=np.uint64)
    C, L = len(q), 96
    st = make_states(C * L, KEY, NONCE, first_counter=12345).reshape(C, L, 16)
    ref = st.copy()
    d, out = b.dev(st), b.empty(C, 4 * L)
    for shift in (0, -1):
        lib.check(lib.tb200_randint_fast(b.device, b.ptr(d), C, L, q.ctypes.data, shift, C * L, b.ptr(out), None),
                  "randint_fast")
        want = oc.randint_fast(ref, [int(v) for v in q], shift, C * L)
        assert np.array_equal(b.host(out), want), "randint_fast samples"
        assert np.array_equal(b.host(d), ref), "randint_fast stepped states"
        got = b.host(out)
        for c in range(C - 1):  # (the 2^64 - 59 channel exceeds int64 by design of the reference's return type)
            assert got[c].min() >= shift and got[c].max() < int(q[c]) + shift
    # two-step variant: in place on random words
    words = oc.chacha20_block(ref)
    w = b.dev(words)
    lib.check(lib.tb200_randint(b.device, b.ptr(w), C, L, q.ctypes.data, None), "randint")
    want = words.copy()
    want[:, :, 0::4] = oc.randint_from_blocks(words, [int(v) for v in q]).reshape(C, L, 4)
    assert np.array_equal(b.host(w), want), "randint in place"


def check_gaussian(lib, b: Buf):
    lut, size, depth = oc.build_cdt_tree()
    rows = 4096
    st = make_states(rows, KEY, NONCE, first_counter=99)
    ref = st.copy()
    d, out = b.dev(st), b.empty(4 * rows)
    lib.check(lib.tb200_discrete_gaussian_fast(b.device, b.ptr(d), rows, lut.ctypes.data, size, depth, rows, b.ptr(out),
                                               None), "discrete_gaussian_fast")
    want = oc.discrete_gaussian_fast(ref, lut, size, depth, rows)
    got = b.host(out)
    assert np.array_equal(got, want), "discrete_gaussian_fast samples"
    assert np.array_equal(b.host(d), ref)
    assert abs(got.mean()) < 0.2 and 2.9 < got.std() < 3.5 and np.abs(got).max() < 32  # sigma = 3.2
    words = oc.chacha20_block(ref)
    w = b.dev(words)
    lib.check(lib.tb200_discrete_gaussian(b.device, b.ptr(w), rows, lut.ctypes.data, size, depth, None),
              "discrete_gaussian")
    want = words.copy()
    want[:, 0::4] = oc.gaussian_from_blocks(words, lut, size, depth).reshape(rows, 4)
    assert np.array_equal(b.host(w), want), "discrete_gaussian in place"


def check_randround(lib, b: Buf):
    rng = np.random.default_rng(5)
    n = 5000
    coef = rng.normal(0, 1e6, n)
    coef[:8] = [0.0, -0.0, 0.5, -0.5, 1.0 - 2.0 ** -33, -3.999999999, 7.0, -7.0]
    words = rng.integers(0, 1 << 32, n, dtype=np.int64)
    words[:8] = [0, 0, (1 << 31) - 1, 1 << 31, 0xFFFFFFFF, 0, 5, 5]
    c, w = b.dev(coef, np.float64), b.dev(words)
    lib.check(lib.tb200_randround(b.device, b.ptr(c), b.ptr(w), n, None), "randround")
    got = b.host(w)
    assert np.array_equal(got, oc.randround(coef, words)), "randround"
    assert np.abs(got - coef).max() <= 1.0


ALL = (check_chacha20, check_randint, check_gaussian, check_randround)
