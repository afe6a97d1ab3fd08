"""TEST INFRASTRUCTURE ONLY: runs bench.py's own arm end to end on a machine without a GPU -- libmifgpu's kernels through
the SIMT interpreter (MIFGPU_LIB), torch.cuda's stream / event / pinning calls replaced by inert stand-ins -- to check
that the JSON line is assembled without a Python error and carries every key of the contract.  The numbers it prints
are meaningless (timings of an interpreter); nothing here is ever used to report performance.
usage: bench_dry_run.py [bench.py arguments]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
assert "simt" in os.environ.get("MIFGPU_LIB", ""), "MIFGPU_LIB must point at the SIMT build"

import torch  # noqa: E402


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = FakeEvent
torch.cuda.ExternalStream = lambda *a, **k: None
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self

# several ranks (torchrun): gloo instead of NCCL for torch.distributed, host tensors instead of device tensors
import torch.distributed as dist  # noqa: E402

_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **kwargs: _init("gloo")
_zeros, _tensor = torch.zeros, torch.tensor
torch.zeros = lambda *a, **k: _zeros(*a, **{key: v for key, v in k.items() if key != "device"})
torch.tensor = lambda *a, **k: _tensor(*a, **{key: v for key, v in k.items() if key != "device"})

import bench  # noqa: E402

sys.argv = ["bench.py"] + sys.argv[1:]
rc = bench.main()
sys.exit(rc)
