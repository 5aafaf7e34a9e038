"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (m3vit_b200/).

Pure-PyTorch stand-in for the third-party package the reference's MoE layer
delegates its arithmetic to: FastMoE (`fmoe`), github.com/laekov/fastmoe,
pinned by the reference at commit 4edeccd (reference README.md:42-50; the
Dockerfile env_setup/Dockerfile.cuda12.1:25-30 installs un-pinned master).
FastMoE is not vendored under /root/reference and is not installable here (no
network, CUDA build), so its *published algorithm* is restated: just enough of
the `fmoe` surface that the reference files

    models/moe/origin/custom_moe_layer.py   models/moe/origin/noisy_gate_vmoe.py
    models/moe/ckpt/custom_moe_layer.py     models/moe/ckpt/noisy_gate_vmoe.py

import and run UNMODIFIED on CPU when `oracle/shim` and `/root/reference` are
on sys.path (see oracle/make_golden.py).  Semantics restated (upstream file):

  fmoe/layers.py     FMoE.__init__/expert_fn/mark_parallel_comm,
                     _fmoe_general_global_forward
  fmoe/functions.py  prepare_forward/count_by_gate, MOEScatter, MOEGather,
                     MOELinear, ensure_comm, Slice, AllGather
  fmoe/linear.py     FMoELinear  (weight [E, out, in], bias [E, out])
  fmoe/gates/        BaseGate (tot_expert = num_expert * world_size), NaiveGate

The one deliberate difference: upstream `assign_pos` orders rows inside an
expert queue with atomics (nondeterministic); here a stable sort is used.  Row
order inside a queue does not change any per-row result.
"""
from . import layers, functions, linear, gates  # noqa: F401
