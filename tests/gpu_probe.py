"""tcgen05 descriptor row-shift probe (see csrc/tc_probe.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from srl_zoo_b200._lib import check, lib, ptr, stream_ptr

out = torch.zeros(128, 64, device="cuda")
for mn in (0, 1):
    for mode in (0, 1):
        for r0 in (0, 1, 2, 3, 5, 7, 8, 9, 13, 64, 77):
            out.zero_()
            check(lib.srlz_probe_desc_shift(ptr(out), r0, mode, mn, stream_ptr()), "probe")
            torch.cuda.synchronize()
            o = out.cpu()
            if mn == 0:
                # expect D[i][0] = (r0+i) % 250, D[i][k] = k
                rows = torch.tensor([(r0 + i) % 250 for i in range(128)], dtype=torch.float32)
                ok_rows = torch.equal(o[:, 0], rows)
                ok_k = all(torch.equal(o[:, k], torch.full((128,), float(k))) for k in range(1, 64))
                got = o[:8, 0].tolist()
            else:
                # expect D[m][n] = image[r0+n][m]: m=0 -> (r0+n)%250 ; m>=1 -> m
                rows = torch.tensor([(r0 + n) % 250 for n in range(64)], dtype=torch.float32)
                ok_rows = torch.equal(o[0, :], rows)
                ok_k = all(torch.equal(o[m, :], torch.full((64,), float(m))) for m in range(1, 64))
                got = o[0, :8].tolist()
            print("mn_major=%d base_offset_mode=%d r0=%2d  rows_ok=%s k_ok=%s  first=%s" % (mn, mode, r0, ok_rows, ok_k, got), flush=True)
