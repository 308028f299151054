"""Real multi-GPU run of the z-slab path (one process per GPU, NCCL), when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_slabs_match_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (the one-device slab emulation is covered by tests/test_slabs.py)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "slab_parity_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("SLAB_PARITY_OK") == 2, out.stdout[-3000:]


@pytest.mark.gpu
def test_tessellate_multi_over_real_devices():
    """mc_tessellate_multi with one context per device (peer copies over NVLink) against one GPU."""
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (several contexts on one device are covered by tests/test_slabs.py)")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import canon
    import moosecad_b200 as mc
    from moosecad_b200 import _lib, scenes, slabs
    world = min(n, 4)
    g = mc.Geom()
    obj = scenes.synthetic_csg(g, 20, seed=3)
    g.backend.set_parameters(0.1, 1.0)
    res = 1.0 / 90
    ctxs = [mc.Context(d) for d in range(world)]
    whole = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctxs[0]).tessellate()
    handles = slabs.tessellate_multi(ctxs, obj, res, 0.2)
    try:
        coords, tets, stats = slabs.gather_slab_meshes(handles)
        canon.assert_same_mesh(canon.canonical(whole.vertices(), whole.elems()), canon.canonical(coords, tets))
        ws = whole.stats()
        for key in ("ntets_stage", "stats", "aspect_ratio", "inverted"):
            assert stats[key] == ws[key], key
        assert np.unique(tets).size == len(coords)
    finally:
        for h in handles:
            _lib.lib().mc_mesh_free(h)
