"""GPU x2: CFG split over a GPU pair through the C API (needs two devices; skipped on one).  The split run is compared with the ORACLE
generate loop, not with the unsplit run."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cfg_split_pair_matches_oracle(tmp_path):
    out = tmp_path / "pair.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "pair_check.py"), "--S", "16", "--images", "3", "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    print(res)
    assert res["replicated_bitwise"]                       # both ranks applied the identical update: no second exchange needed
    assert res["decode_split"] == [0, 2]                   # rank 0 decodes ceil(3/2) images
    assert res["psnr_vs_oracle_db"] >= 35.0 and res["final_latent_rel_l2_vs_oracle"] < 5e-2
    assert res["split_vs_unsplit_latent_rel_l2"] < 2e-2    # same kernels, batch 3 vs batch 6 plans
