"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares, the ctypes
struct layout equals the C compiler's, argument validation returns error codes, and the product refuses CPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpde_b200.h")


@pytest.fixture(scope="module")
def ffi():
    from dynamical_pde_diffusion_b200 import _build, _ffi

    if not os.path.exists(_ffi.LIB_PATH):
        _build.build_library()
    return _ffi


def test_library_exports_every_declared_symbol(ffi):
    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(dpde_\w+)\s*\(", text, flags=re.M))
    assert declared == set(ffi.EXPORTED), declared ^ set(ffi.EXPORTED)
    L = ffi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.dpde_abi_version() == ffi.ABI_VERSION == 2
    assert L.dpde_guidance_workspace_bytes() >= 3 * 8 * 148


def test_struct_layout_matches_the_c_compiler(ffi, tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dpde_b200.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dpde_view), sizeof(dpde_guidance_desc),'
                   'offsetof(dpde_guidance_desc, x0), offsetof(dpde_guidance_desc, mask_u), offsetof(dpde_guidance_desc, sample_coef),'
                   'offsetof(dpde_guidance_desc, dx), offsetof(dpde_guidance_desc, gamma), offsetof(dpde_guidance_desc, easy_axis));'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(dpde_mailbox), offsetof(dpde_mailbox, epoch), offsetof(dpde_mailbox, boxes),'
                   'sizeof(dpde_halo_peers), offsetof(dpde_halo_peers, ticket), offsetof(dpde_halo_peers, epoch), offsetof(dpde_halo_peers, H_down),'
                   'DPDE_MAX_RANKS, DPDE_MAILBOX_BYTES);return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    D = ffi.GuidanceDesc
    M, P = ffi.Mailbox, ffi.HaloPeers
    want = [C.sizeof(ffi.View), C.sizeof(D), D.x0.offset, D.mask_u.offset, D.sample_coef.offset, D.dx.offset, D.gamma.offset,
            D.easy_axis.offset,
            C.sizeof(M), M.epoch.offset, M.boxes.offset, C.sizeof(P), P.ticket.offset, P.epoch.offset, P.H_down.offset,
            ffi.MAX_RANKS, ffi.MAILBOX_BYTES]
    assert got == want


def test_argument_validation_without_a_gpu(ffi):
    L = ffi.lib()
    assert L.dpde_euler_predict(None, None, 1.0, 0.5, None, 4, None) == -1
    assert b"null" in L.dpde_last_error()
    assert L.dpde_heun_guided_update(None, None, None, None, None, 1.0, 0.5, None, None, 4, None) == -1
    assert L.dpde_laplacian(None, None, 0, 1, 8, 8, 64, 0.1, 0, None) == -1
    d = ffi.GuidanceDesc()
    d.B, d.C, d.ch_a, d.H, d.W, d.pde_kind = 1, 2, 1, 1, 8, ffi.PDE_HEAT          # H = 1 cannot be reflect-padded
    assert L.dpde_guidance_vjp(C.byref(d), None, None, None, None, None) == -1
    assert b"H and W" in L.dpde_last_error()
    d.H, d.pde_kind = 8, 99
    d.x0.ptr = 1
    assert L.dpde_guidance_reduce(C.byref(d), None, None, 0, None, None, None) == -1
    d.pde_kind, d.C = ffi.PDE_LLG_NORM, 5                                          # LLG kinds need 3 u-channels
    assert L.dpde_guidance_reduce(C.byref(d), None, None, 0, None, None, None) == -1
    assert b"3 magnetisation" in L.dpde_last_error()
    with pytest.raises(ffi.DpdeError):
        ffi.call("dpde_halo_pack", None, 0, 1, 8, 8, 2, None, None, None)
    # mailbox exchange and the fused update + push: bad descriptors never reach a launch
    d.pde_kind, d.C, d.sample_coef, d.dx = ffi.PDE_HEAT, 2, 1, 0.1
    mb = ffi.Mailbox()
    mb.world, mb.rank, mb.epoch = 9, 0, 1
    assert L.dpde_guidance_reduce_post(C.byref(d), 1, 1, C.byref(mb), None) == -1 and b"world" in L.dpde_last_error()
    mb.world, mb.epoch = 2, 0
    assert L.dpde_mailbox_wait_finalize(C.byref(d), C.byref(mb), 1.0, None, 1, 1, None, None) == -1 and b"epoch" in L.dpde_last_error()
    mb.epoch = 1
    assert L.dpde_mailbox_wait_finalize(C.byref(d), C.byref(mb), 1.0, None, 1, 1, None, None) == -1 and b"boxes[0]" in L.dpde_last_error()
    hp = ffi.HaloPeers()
    assert L.dpde_heun_guided_update_rows_push(1, 1, None, None, None, 1.0, 0.0, 1, 1, 2, 7, 8, 2, C.byref(hp), None) == -1
    assert b"4 halo" in L.dpde_last_error()
    assert L.dpde_heun_guided_update_rows_push(1, 1, None, None, None, 1.0, 0.0, 1, 1, 2, 8, 8, 2, C.byref(hp), None) == -1
    assert b"ticket" in L.dpde_last_error()


def test_product_has_no_cpu_path():
    import dynamical_pde_diffusion_b200 as dp

    u = torch.zeros(2, 1, 8, 8, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA"):
        dp.laplacian(u, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        dp.heat_loss2(u, u, torch.ones(2, 2), 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        dp.llg_loss2(torch.zeros(2, 3, 8, 8), None, None)
    smp = dp.JointSampler(lambda *a: a[0], torch.device("cpu"), (8, 8), 2, 2, 1, dp.heat_loss2, {"dx": 0.1}, num_steps=3)
    z = torch.zeros(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        smp.sample(torch.ones(2, 2), z, z, z > 1, z > 1, 1.0, 1.0, 1.0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import it or the reference."""
    pkg = os.path.join(ROOT, "dynamical_pde_diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f
