"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone) -- CPU restatement of the reference's T-map consumer:

  load_typicality          /root/reference/diffmining/typicality/cluster.py:125-137
  pool                     /root/reference/diffmining/typicality/utils.py:74-80
  df_D (window table)      /root/reference/diffmining/typicality/cluster.py:183-205
  get_non_overlapping      /root/reference/diffmining/typicality/utils.py:94-102

`patch_scores` follows the reference's own order of operations on the raw fp16 loss grid [N, 2, 4, h, w]
(index 0 = condition, 1 = unconditional, as compute.py:187-188 stacks them); `non_overlapping_topk` is the pandas
loop of get_non_overlapping re-expressed on arrays (same inclusive overlap test, same "take the first row of the
sorted frame" rule)."""
import numpy as np
import torch
import torch.nn.functional as F


def patch_scores(grid_f16: np.ndarray, H: int, W: int, kx: int, ky: int) -> np.ndarray:
    dm = torch.from_numpy(np.asarray(grid_f16)).float()              # cluster.py:129-131
    dm = dm.mean(dim=2)                                              # channel mean       :132
    dm = F.interpolate(dm, (H, W), mode="bilinear")                  # resize             :134
    pool = torch.nn.AvgPool2d((kx, ky), stride=(1, 1), padding=0)    # utils.py:74-80
    d = pool(dm[:, 0].unsqueeze(1)) - pool(dm[:, 1].unsqueeze(1))    # cluster.py:135
    return (-d.squeeze(1).mean(dim=0)).numpy()                       # :136


def non_overlapping_topk(D: np.ndarray, kx: int, ky: int, k: int, ascending: bool = False):
    """rows (x_start, y_start, x_end, y_end, D) like df_D + get_non_overlapping; ties: lowest (i, j) first"""
    Ho, Wo = D.shape
    order = np.argsort(D.ravel() if ascending else -D.ravel(), kind="stable")
    alive = np.ones(Ho * Wo, dtype=bool)
    ii, jj = np.divmod(np.arange(Ho * Wo), Wo)
    out = []
    ptr = 0
    while len(out) < k:
        while ptr < order.size and not alive[order[ptr]]:
            ptr += 1
        if ptr >= order.size:
            break
        idx = order[ptr]
        i, j = int(ii[idx]), int(jj[idx])
        out.append((i, j, i + kx, j + ky, float(D[i, j])))
        # utils.py:98: drop rows with x_start <= x_end* and x_end >= x_start* and y_start <= y_end* and y_end >= y_start*
        alive &= ~((ii <= i + kx) & (ii + kx >= i) & (jj <= j + ky) & (jj + ky >= j))
    return out
