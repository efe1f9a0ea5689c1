"""CPU suite: the C-ABI library builds, loads, and exports every symbol include/svr_b200.h declares
(no compute calls here -- there is no GPU in the build container)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _declared():
    text = (REPO / "include" / "svr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import svr_b200
    from svr_b200 import _abi
    lib = _abi.load()
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(str(_abi.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/svr_b200.h but not exported"
    assert set(names) == set(_abi.EXPORTS), set(names) ^ set(_abi.EXPORTS)
    assert lib.svr_abi_version() == 1


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    import svr_b200
    proj = svr_b200.project((16, 16, 16), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5]))
    with pytest.raises(RuntimeError, match="CUDA"):
        proj.pc_voxels(torch.zeros(1, 4, 3))
    svr_b200.configure(net_res=128)
    net = svr_b200.IFNet().eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 1, 16, 16, 16), torch.zeros(1, 4, 3))


def test_product_does_not_import_oracle():
    pkg = REPO / "single-view-3d-reconstruction_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "svr_oracle" not in src, f


def test_state_dict_keys_match_reference():
    import svr_b200
    from oracle import ref_torch as R
    for net_res in (128, 32):
        svr_b200.configure(net_res=net_res)
        net = svr_b200.IFNet()
        mine = {k: tuple(v.shape) for k, v in net.state_dict().items() if not k.endswith("num_batches_tracked")}
        assert mine == R.ifnet_param_shapes(net_res)
    svr_b200.configure(net_res=128)
    import torch
    proj = svr_b200.project((139, 104, 112), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5]))
    assert list(proj.state_dict().keys()) == ["sigma"]


def test_get_intrinsic_parses_the_reference_file_format(tmp_path):
    """projection.py:209-218 reads focal length, cx, cy from the first two rows of a text matrix in the format of the
    reference's data/raw/overfit/00000/intrinsic.txt; without a file the package falls back to the same constants."""
    import torch
    import svr_b200
    f = tmp_path / "intrinsic.txt"
    f.write_text("[[300.25,   0.       , 161.5,  0.],\n[  0.       , 300.25, 118.25,  0.],\n"
                 "[  0.       ,   0.       ,   1. ,  0.],\n[  0.       ,   0.       ,   0. ,  1.]]")
    K = svr_b200.project.get_intrinsic(f)
    want = torch.tensor([[300.25, 0, 161.5, 0], [0, 300.25, 118.25, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    assert K.dtype == torch.float32 and torch.equal(K, want)
    K0 = svr_b200.project.get_intrinsic(tmp_path / "missing.txt")
    assert abs(float(K0[0, 0]) - 277.1281435) < 1e-4 and float(K0[0, 2]) == 159.5 and float(K0[1, 2]) == 119.5


def test_precision_switch_controls_tf32_and_restores_it():
    import torch
    import svr_b200
    before = torch.backends.cudnn.allow_tf32
    svr_b200.configure(precision=32)
    assert torch.backends.cudnn.allow_tf32 is False
    svr_b200.configure(precision=16)
    assert torch.backends.cudnn.allow_tf32 == before


@pytest.mark.parametrize("ks", [(3, 3, 3), (5, 5, 5), (5, 3, 7)])
def test_smoothing_kernel_matches_reference_recipe(ks):
    """project.smoothing_kernel (projection.py:82-100): three normalised 1-D Gaussians of sigma.  Equal tap counts take the
    batched (3, k) evaluation, mixed ones the per-axis loop: taps and d(taps)/d(sigma) must equal the oracle's restatement
    of the reference recipe bit for bit (pure torch ops, so this runs on the CPU)."""
    import torch
    import svr_b200
    from oracle import ref_torch as R
    sig = torch.tensor([1.5, 0.9, 2.1])
    proj = svr_b200.project((16, 12, 20), list(ks), sig.clone())
    mine = proj.smoothing_kernel()
    s_ref = sig.clone().requires_grad_(True)
    want = R.smoothing_kernels(s_ref, list(ks))
    assert [tuple(t.shape) for t in mine] == [tuple(t.shape) for t in want]
    cot = [torch.linspace(0.3, 1.7, k) for k in ks]
    for a in range(3):
        assert torch.equal(mine[a].detach(), want[a].detach()), a
    sum((m.reshape(-1) * c).sum() for m, c in zip(mine, cot)).backward()
    sum((w.reshape(-1) * c).sum() for w, c in zip(want, cot)).backward()
    torch.testing.assert_close(proj.sigma.grad, s_ref.grad, rtol=1e-6, atol=1e-7)


def test_voxeliser_ownership_masks_match_their_definition():
    """csrc/projection.cu hard-codes two constant tables for the accumulate pass (neighbourhood bit ((dz+1)*3 + (dy+1))*3 +
    (dx+1) of the source cell of pass ps for the voxel at corner shift sh).  Re-derive them from the definition
    (projection.py:69-78: pass ps adds to the voxel at floor + ps, so the voxel cell+sh receives pass ps from the cell
    cell + sh - ps) and check the magic-number divisions by 9 and 3 that decode a bit index."""
    src = (REPO / "single-view-3d-reconstruction_b200" / "csrc" / "projection.cu").read_text()

    def table(name):
        m = re.search(name + r"\[8\]\s*=\s*\{([^}]*)\}", src)
        assert m, name
        return [int(v.strip().rstrip("u"), 16) for v in m.group(1).split(",")]

    lower, srcs = [], []
    for sh in range(8):
        lo = sr = 0
        for ps in range(8):
            n = [((sh >> k) & 1) - ((ps >> k) & 1) for k in (2, 1, 0)]          # (dz, dy, dx) of the source cell
            bit = ((n[0] + 1) * 3 + (n[1] + 1)) * 3 + (n[2] + 1)
            if ps < sh:
                lo |= 1 << bit
            else:
                sr |= 1 << bit
        lower.append(lo)
        srcs.append(sr)
    assert table("c_vox_lower") == lower
    assert table("c_vox_src") == srcs
    for bit in range(27):
        q9 = (bit * 57) >> 9
        r9 = bit - q9 * 9
        q3 = (r9 * 11) >> 5
        assert (q9, r9, q3) == (bit // 9, bit % 9, r9 // 3)
    # a higher bit is an earlier pass for every shift: walking the set bits downwards is the reference's pass order
    for sh in range(8):
        order = []
        for bit in range(26, -1, -1):
            if (srcs[sh] >> bit) & 1:
                nz, ny, nx = bit // 9 - 1, (bit % 9) // 3 - 1, bit % 3 - 1
                order.append(((((sh >> 2) & 1) - nz) << 2) | ((((sh >> 1) & 1) - ny) << 1) | ((sh & 1) - nx))
        assert order == sorted(order) and order[0] == sh
