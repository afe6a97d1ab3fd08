"""TEST INFRASTRUCTURE ONLY: the asynchronous-transfer API of libmifgpu under the SIMT interpreter with
MIF_EMU_LAZY_COPIES set by the caller (tests/test_simt_emu.py): asynchronous copies, event records and stream waits are
then queued per stream and carried out as late as the programming model allows (or, with "odd" / "even", every second
stream races ahead), so that a missing event dependency or a staging buffer refilled too early changes the result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    assert "simt" in os.environ.get("MIFGPU_LIB", "") and os.environ.get("MIF_EMU_LAZY_COPIES")
    import mif_b200 as mif
    import test_gpu_vs_oracle as parity
    # a new tensor is usable by an asynchronous upload at once (its zero fill is ordered before the transfer), six tensors
    # in flight through the two staging buffers of each direction, no compute call in between
    ctx, grid = parity.make_pair(mif, (12, 9, 8), (False, False, False))
    rng = np.random.default_rng(5)
    fields = [rng.uniform(-1, 1, grid.shape(c % 4)) for c in range(6)]
    tensors = [ctx.tensor(c % 4) for c in range(6)]
    got = [np.empty_like(h) for h in fields]
    for t, h in zip(tensors, fields):
        t.upload_async(h)
    for t, out in zip(tensors, got):
        t.download_async(out)
    ctx.synchronize()
    assert all(np.array_equal(a, b) for a, b in zip(got, fields))
    ctx.close()
    # three jobs rotated through two device field sets against the blocking calls (the GPU test's own body, on a grid the
    # interpreter steps through in seconds)
    parity.test_async_transfers_pipeline_matches_synchronous_calls(mif, N=(12, 9, 8))
    parity.test_timestep_random_state(mif, (9, 7, 6), (False, False, False), "ethier_steinman")
    print("async cases ok")


if __name__ == "__main__":
    main()
