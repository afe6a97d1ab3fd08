"""TEST INFRASTRUCTURE ONLY: runs a few bodies of the GPU parity tests with libmifgpu's kernels executed by the SIMT
interpreter (MIFGPU_LIB must point at tests/simt_emu/build/libmifgpu_simt.so).  Called by tests/test_simt_emu.py in a
subprocess, because the library a process binds is fixed at first use."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    assert "simt" in os.environ.get("MIFGPU_LIB", ""), "MIFGPU_LIB must point at the SIMT build"
    import mif_b200 as mif
    assert "simt" in mif.LIB_PATH
    import __graft_entry__ as entry
    import test_gpu_golden as golden
    import test_gpu_vs_oracle as parity

    F, T = False, True
    # `--shard i/n` runs every n-th case starting at i, so that the caller can spread the list over processes
    shard, shards = (int(v) for v in sys.argv[sys.argv.index("--shard") + 1].split("/")) if "--shard" in sys.argv else (0, 1)
    solve = parity.test_pressure_solve_random_velocity
    cases = [lambda: entry.smoke()]                                   # 16^3, generic sweeps, stage / bc / correct kernels
    for N, periodic in [((513, 4, 9), (F, F, F)),                     # x sweeps on the 16 x 32 transform (M = 512)
                        ((9, 257, 3), (F, F, F)),                     # TMA-staged strided y sweeps, radix-8 passes (M = 256)
                        ((11, 3, 513), (F, F, F)),                    # TMA-staged fused z sweep on the 16 x 32 transform
                        ((1025, 3, 4), (F, F, F)),                    # x sweeps of 1025-point lines (two 512-point halves)
                        ((20, 9, 513), (F, F, T)),                    # warp real-FFT path, fused z sweep
                        ((65, 9, 65), (F, F, F)),                     # CTA-synchronous fast path
                        ((12, 10, 14), (F, F, T)),                    # Bluestein
                        ((10, 1025, 3), (F, F, F))]:                  # TMA-staged 1025-point lines (split transform), two x tiles
        cases.append(lambda N=N, periodic=periodic: solve(mif, N, periodic))
    cases.append(lambda: parity.test_async_transfers_pipeline_matches_synchronous_calls(mif))  # asynchronous transfer API
    cases.append(lambda: parity.test_timestep_random_state(mif, (9, 10, 12), (T, T, T), "test_case_2"))
    cases.append(lambda: parity.test_timestep_nhn_matches_oracle(mif))
    cases.append(lambda: golden.test_timestep_velocity_matches_reference(mif, "vtest_12_2"))
    for index, case in enumerate(cases):
        if index % shards == shard:
            case()
            print("case", index, "ok", flush=True)
    print("simt cases ok")


if __name__ == "__main__":
    main()
