"""Diagnosis aid for tests/test_gpu_zz_head_mirrors.py::test_triplet_head_golden: is a gradient mismatch against the
reference's float64 golden a bilinear-cell knife edge (a sampling coordinate within float32 noise of an integer, where
d out/dH is one-sided) or a real discrepancy?  Runs the head on cuda:0 in float32 and, through the oracle's closed
forms on the CPU in float64, once with floor() taken in float64 and once with the cell fixed to the kernel's floor().
Test infrastructure (imports oracle/ and tests/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import cpu_kernels  # noqa: E402
from test_gpu_kernels import _kernel_cells, _warp_direct_autograd  # noqa: E402
from test_head_mirrors import TRIPLET_CASES, triplet_forward  # noqa: E402


class Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def main(name='aware'):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'triplet_head_P32.npz')))
    loss, g12, g21, gnorm, conv1 = triplet_forward(g, name, torch.float32, 'cuda')
    g12 = g12.cpu().double().numpy()
    ref = g[name + '_g12_64']
    print('loss gpu %.9g golden64 %.9g golden32 %.9g' % (loss.item(), float(g[name + '_loss64']), float(g[name + '_loss32'])))
    for b in range(ref.shape[0]):
        print('sample %d: g12 rel err vs golden64 %.3e   (golden32 vs golden64 %.3e)' % (
            b, np.linalg.norm(g12[b] - ref[b]) / np.linalg.norm(ref[b]),
            np.linalg.norm(g[name + '_g12_32'][b] - ref[b]) / np.linalg.norm(ref[b])))
    if g21 is not None:
        r21 = g[name + '_g21_64']
        for b in range(r21.shape[0]):
            print('sample %d: g21 rel err vs golden64 %.3e' % (b, np.linalg.norm(g21[b].cpu().numpy() - r21[b]) / np.linalg.norm(r21[b])))

    # the H the device computed, and the coordinates closest to an integer
    import bihome_b200.functional as F
    d = torch.cat([torch.as_tensor(g['delta_12']), torch.as_tensor(g['delta_21'])]).float().cuda()
    H32 = F.dlt4(d, size=(32, 32)).cpu().numpy()
    h = H32.reshape(-1, 9).astype(np.float64)[:, :, None, None]
    ys, xs = np.meshgrid(np.arange(32.0), np.arange(32.0), indexing='ij')
    w = h[:, 6] * xs + h[:, 7] * ys + h[:, 8]
    u = (h[:, 0] * xs + h[:, 1] * ys + h[:, 2]) / w
    v = (h[:, 3] * xs + h[:, 4] * ys + h[:, 5]) / w
    cu_, cv_ = _kernel_cells(H32, 32, 32)
    for b in range(u.shape[0]):
        du = np.abs(u[b] - np.round(u[b]))
        dv = np.abs(v[b] - np.round(v[b]))
        flips = int((np.floor(u[b]) != cu_[b].numpy()).sum() + (np.floor(v[b]) != cv_[b].numpy()).sum())
        print('H %d: min |u-int| %.3e at %s, min |v-int| %.3e at %s, kernel-vs-float64 cell flips %d' % (
            b, du.min(), np.unravel_index(du.argmin(), du.shape), dv.min(), np.unravel_index(dv.argmin(), dv.shape), flips))

    # CPU float64 through the oracle with the cells fixed to the kernel's
    F = cpu_kernels.install(Patch())

    def warp_fixed(src, H, out_h, out_w, pool=None):
        cells = _kernel_cells(H.detach().numpy().astype(np.float32), out_h, out_w)
        out = _warp_direct_autograd(src, H, out_h, out_w, cells)
        if pool:
            return out, cpu_kernels.coverage_mask(H, src.shape[-2:], (out_h, out_w), pool)
        return out
    F.warp = warp_fixed
    _, f12, f21, _, _ = triplet_forward(g, name, torch.float64, 'cpu')
    f12 = f12.numpy()
    for b in range(ref.shape[0]):
        print('sample %d: g12 gpu vs fixed-cell float64 %.3e ; fixed-cell float64 vs golden64 %.3e' % (
            b, np.linalg.norm(g12[b] - f12[b]) / np.linalg.norm(f12[b]), np.linalg.norm(f12[b] - ref[b]) / np.linalg.norm(ref[b])))


if __name__ == '__main__':
    for n in (sys.argv[1:] or ['aware']):
        print('====', n)
        main(n)
