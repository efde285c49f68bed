"""Drives the host-compiled kernels (tests/emu) built with -fsanitize=address or =thread through a small but
complete workload (every kernel class, toy rings), each result checked against the oracle.
AddressSanitizer sees every out-of-bounds global / "shared" access of the kernels' index arithmetic;
ThreadSanitizer sees every pair of conflicting shared-memory accesses not separated by a barrier
(one std::thread per CUDA thread, __syncthreads = std::barrier).  compute-sanitizer itself is closed on
the GPU pool (gpurun refuses it), so this is the memcheck / racecheck evidence of the tile kernels.

  sh tests/emu/run_sanitizers.sh        # builds both variants, runs this file under each, logs to profiles/
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.path.insert(0, os.path.dirname(TESTS))


def main():
    import parity
    from parity import Harness, Setup
    from tiberate_fhe_b200 import _native

    lib = _native.Lib(sys.argv[1])
    h = Harness(lib, use_torch=False)
    for logN, ns, K in ((8, 3, 2), (12, 2, 2)):
        s = Setup.toy(h, logN, ns, K, seed=logN, rot_deltas=(1,))
        try:
            parity.check_ntt(s, 0, True, 2)
            parity.check_pointwise(s, 0, True)
            parity.check_he_ops(s, 0)
            for mode in ((False, 0), (True, 0), (True, 8, 0, 1), (True, 8)):
                parity.set_mode(s, mode)
                parity.check_engine(s, 0, ops=("rescale", "keyswitch", "rotate", "cc_mult", "pc_mult"))
            if logN == 8:
                s.ctx.set_chunk(2)
                parity.check_engine(s, 1, batch=3, ops=("cc_mult", "rotate"))
        finally:
            s.close()
        print("ok", logN, flush=True)


if __name__ == "__main__":
    main()
