"""CPU-side tests: the C-ABI library loads and exports every symbol include/rtpose_b200.h declares (no compute calls),
host logic (tap lists, target assignment, config loading, registry/builder, state_dict names) and the
no-CPU-fallback contract."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CFG = "/root/reference/configs/cruw_pose"


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rtpose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rtpose_b200 import build, lib
    if not os.path.exists(lib.LIB_PATH):
        build.build()
    L = lib.load()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "librtpose_b200.so does not export %s" % n
    missing = [n for n in names if n not in lib.PROTOTYPES]
    assert not missing, "ctypes prototypes missing for %s" % missing
    extra = [n for n in lib.PROTOTYPES if n not in names]
    assert not extra, "prototypes not declared in the header: %s" % extra
    assert L.rtp_version() >= 100
    assert isinstance(L.rtp_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu():
    """Argument validation runs before any CUDA call, so it can be exercised on the CPU box."""
    import ctypes as C
    from rtpose_b200 import lib
    L = lib.load()
    d = lib.ConvDesc()  # all-null descriptor
    assert L.rtp_conv(C.byref(d), None) == -1
    assert b"null" in L.rtp_last_error() or b"is null" in L.rtp_last_error()
    assert L.rtp_conv(None, None) == -1
    assert L.rtp_scale_f32(None, 8, 1.0, None) == -1
    assert L.rtp_conv_k3s1_smem_bytes(32, 32, 16, 160, 64) > 0
    assert L.rtp_conv_k3s1_smem_bytes(32, 128, 16, 160, 64) == -1      # 3*NPo > 256
    assert L.rtp_wgrad_k3s1_supported(32, 32, 16, 160, 64) == 1
    assert L.rtp_wgrad_k3s1_supported(64, 32, 16, 160, 64) == 0


def test_no_cpu_fallback():
    from rtpose_b200 import det3d_compat as D
    from rtpose_b200 import lib
    m = D.HRNet3D(backbone_cfg="hr_tiny_feat32_zyx_l4_in32", final_conv_in=192, final_conv_out=128, final_fuse="conat_conv", ds_factor=1)
    with pytest.raises(lib.RtpError):
        m(torch.zeros(1, 32, 8, 16, 16))  # CPU tensor -> loud failure, never a PyTorch-op fallback
    if not torch.cuda.is_available():
        with pytest.raises(lib.RtpError):
            lib.require_device()


def test_tap_lists():
    from rtpose_b200 import ops
    f = ops.taps_fwd(3)
    assert len(f) == 27 and f[0] == (-1, -1, -1, 0) and f[13] == (0, 0, 0, 13)
    # forward tap (kz,ky,kx) -> (tz, tx, ty): x is the slow in-plane axis of the P8 layout
    assert f[(0 * 3 + 2) * 3 + 1] == (-1, 0, 1, 7)
    d = ops.taps_dgrad_s1(3)
    assert all(a[0] == -b[0] and a[1] == -b[1] and a[2] == -b[2] and a[3] == b[3] for a, b in zip(d, f))
    # stride-2 dgrad: the 8 parity classes partition the 27 taps
    seen = []
    for pz in range(2):
        for px in range(2):
            for py in range(2):
                t = ops.taps_dgrad_s2(pz, px, py)
                assert len(t) == (1 if pz == 0 else 2) * (1 if px == 0 else 2) * (1 if py == 0 else 2)
                seen += [w for _, _, _, w in t]
    assert sorted(seen) == list(range(27))
    # 1-D check of the index relation i = 2*o + k - 1
    for p in range(2):
        for k, t in ([(1, 0)] if p == 0 else [(0, 1), (2, 0)]):
            for ih in range(4):
                i = 2 * ih + p
                o = ih + t
                assert i == 2 * o + k - 1


@pytest.mark.parametrize("one_hm", [True, False])
def test_targets_match_oracle(one_hm):
    from rtpose_b200 import targets
    grid = (16, 64, 160)
    rs = np.random.RandomState(3)
    poses = targets.random_poses(rs, 5, grid)
    poses[1, 3] = [-50.0, 0.0, 0.0]  # joint outside the ROI -> skipped in the 15-class assigner
    got = targets.assign(poses, grid, one_hm, min_radius=2 if one_hm else 1)
    ref = O.batch_targets(list(poses), grid, one_hm)
    for k in ("hm", "ind", "mask", "cat", "anno_pose"):
        np.testing.assert_array_equal(got[k], ref[k].numpy(), err_msg=k)
    assert got["hm"].max() == 1.0


@pytest.mark.parametrize("cfg", sorted(O.CONFIGS))
def test_state_dict_names_match_reference_spec(cfg):
    from rtpose_b200 import det3d_compat as D
    c = O.CONFIGS[cfg]
    names = G_POSE[:c["ncls"]]
    model_cfg = dict(type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
                     backbone=dict(type="HRNet3D", backbone_cfg=c["arch"], final_conv_in=c["final_in"], final_conv_out=c["final_out"],
                                   final_fuse=c["fuse"], ds_factor=1),
                     pose_head=dict(type="CenterHead", tasks=[dict(num_class=c["ncls"], class_names=names)], in_channels=c["head_in"],
                                    share_conv_channel=c["share"], dataset="cruw_pose", weight=c["weight"], code_weights=c["code_weights"],
                                    common_heads={"reg": (c["reg"], 2)}, dcn_head=False),
                     neck=None)
    m = D.build_detector(model_cfg, train_cfg=None, test_cfg=None)
    sd = m.state_dict()
    spec = dict(O.state_dict_spec(cfg))  # pinned against the reference by load_state_dict(strict=True) in make_golden
    assert set(sd) == set(spec)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    # reference init conventions that matter: hm bias -2.19, GroupNorm gamma 1 / beta 0
    assert torch.allclose(sd["pose_head.tasks.0.hm.2.bias"], torch.full_like(sd["pose_head.tasks.0.hm.2.bias"], -2.19))
    assert float(sd["backbone.backbone.layer1.conv2.groupnorm.weight"].min()) == 1.0
    m.load_state_dict(O.synth_state_dict(cfg), strict=True)


G_POSE = ["Pelvis", "Right_Hip", "Right_Knee", "Right_Ankle", "Left_Hip", "Left_Knee", "Left_Ankle", "Thomx", "Head",
          "Left_Shoulder", "Left_Elbow", "Left_Wrist", "Right_Shoulder", "Right_Elbow", "Right_Wrist"]


def test_registry_and_builder_errors():
    from rtpose_b200 import det3d_compat as D
    with pytest.raises(KeyError):
        D.build_detector(dict(type="NoSuchNet"))
    with pytest.raises(KeyError):
        D.build_backbone(dict(type="HRNet3D", backbone_cfg="nope", final_conv_in=1, final_conv_out=1, final_fuse="top"))
    with pytest.raises(TypeError):  # the reference's dcn_head=True is a TypeError too (center_head.py:152)
        D.build_head(dict(type="CenterHead", in_channels=32, tasks=[dict(num_class=1, class_names=["Pelvis"])],
                          share_conv_channel=32, common_heads={"reg": (45, 2)}, dcn_head=True))
    with pytest.raises(KeyError):
        D.DETECTORS.register_module(D.RadarPoseNet)  # already registered
    assert set(["RadarPoseNet"]) <= set(D.DETECTORS.module_dict)
    assert D.install_as_det3d().build_detector is D.build_detector
    from det3d.ops.dcn import DeformConv, ModulatedDeformConvPack  # center_head.py:18; det3d/ops/dcn/__init__.py
    from det3d.ops.dcn.deform_conv import deform_conv
    from rtpose_b200 import dcn
    assert DeformConv is dcn.DeformConv and ModulatedDeformConvPack is dcn.ModulatedDeformConvPack and deform_conv is dcn.deform_conv


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference configs not present (GPU box)")
@pytest.mark.parametrize("name", ["hr3d", "hr3d_one_hm", "hr3d_one_hm_doppler", "hr3d_one_hm_doppler_phase"])
def test_reference_configs_load_unchanged_and_build(name):
    from rtpose_b200 import det3d_compat as D
    from rtpose_b200.config import Config
    cfg = Config.fromfile(os.path.join(REF_CFG, name + ".py"))
    assert cfg.model.type == "RadarPoseNet" and cfg.test_cfg.voxel_size == [0.0453125, 0.15703125, 0.3625]
    m = D.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    want = {"hr3d": 2002194, "hr3d_one_hm": 2217006, "hr3d_one_hm_doppler": 2216942, "hr3d_one_hm_doppler_phase": 8297518}
    assert sum(p.numel() for p in m.parameters()) == want[name]
    with pytest.raises(AttributeError):
        cfg.no_such_key


def test_shard_frames_partitions_exactly():
    from rtpose_b200.dist import shard_frames
    for total in (0, 1, 7, 16, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_frames(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_dcn_pack_modules_rename_pre_v2_checkpoint_keys():
    """deform_conv.py:294-323 / :418-446: checkpoints older than module version 2 name the predictor `<name>_offset`;
    the module's _load_from_state_dict hook moves those entries to `<name>.conv_offset` (hook called the way
    nn.Module.load_state_dict calls it, with the whole checkpoint dict and the module's prefix)."""
    import torch
    from rtpose_b200.dcn import DeformConvPack, ModulatedDeformConvPack
    for cls, ch in ((DeformConvPack, 18), (ModulatedDeformConvPack, 27)):
        m = cls(4, 4, 3, padding=1, deformable_groups=1)
        sd = {"conv2." + k: v.clone() for k, v in m.state_dict().items() if not k.startswith("conv_offset")}
        sd["conv2_offset.weight"] = torch.full((ch, 4, 3, 3), 0.25)
        sd["conv2_offset.bias"] = torch.full((ch,), -1.0)
        missing, unexpected, errors = [], [], []
        m._load_from_state_dict(sd, "conv2.", {}, True, missing, unexpected, errors)  # {} = no version metadata -> < 2
        assert "conv2.conv_offset.weight" in sd and "conv2_offset.weight" not in sd and not errors and not missing
        m.conv_offset._load_from_state_dict(sd, "conv2.conv_offset.", {}, True, missing, unexpected, errors)
        assert float(m.conv_offset.weight.mean()) == 0.25 and float(m.conv_offset.bias.mean()) == -1.0 and not errors
        keep = dict(sd)
        m._load_from_state_dict(sd, "conv2.", {"version": 2}, True, missing, unexpected, errors)  # current version: untouched
        assert sd.keys() == keep.keys()


def test_dcn_pack_cache_is_keyed_on_tensor_identity_and_version():
    """The tensor-core DCN path caches bf16 weight packs; a hit needs the same tensor OBJECTS at the same version."""
    import torch
    from rtpose_b200 import dcn
    a, c = torch.zeros(3), torch.zeros(3)
    k = dcn._tensor_key((a, None))
    assert dcn._same_tensors(k, (a, None))
    assert not dcn._same_tensors(k, (c, None)) and not dcn._same_tensors(k, (a, c)) and not dcn._same_tensors(None, (a, None))
    a.add_(1)  # in-place update (an optimizer step) bumps the version counter
    assert not dcn._same_tensors(k, (a, None))
    assert dcn.tc_supported(128, 128, 3, 3, 4) and not dcn.tc_supported(8, 8, 3, 3, 4) and not dcn.tc_supported(128, 512, 3, 3, 4)
    assert dcn._tc_chunk(256, 128, 9, 64, 160, 4) == dcn.TC_SAMPLE_BYTES // (16 * 9 * 162 * 66 * 16)


def test_one_cycle_schedule_equals_reference_golden():
    """rtpose_b200.optim.one_cycle against the values the reference's own OneCycle class produced
    (det3d/solver/learning_schedules_fastai.py:53-95; oracle/make_sched_golden.py -> tests/golden/one_cycle_golden.json),
    every step of three schedules, relative 1e-12 (math.cos vs numpy.cos)."""
    import json
    import os
    from rtpose_b200.optim import one_cycle
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "one_cycle_golden.json")))
    assert [c["total_step"] for c in cases] == [1000, 37, 10]
    for c in cases:
        for step, (lr, mom) in enumerate(c["lr_mom"]):
            got = one_cycle(step, c["total_step"], lr_max=c["lr_max"], div_factor=c["div_factor"], pct_start=c["pct_start"],
                            moms=tuple(c["moms"]))
            assert abs(got[0] - lr) <= 1e-12 * abs(lr) and abs(got[1] - mom) <= 1e-12 * abs(mom), (c["total_step"], step, got, lr, mom)


def test_header_is_plain_c_and_a_c_program_can_bind_the_library(tmp_path):
    """include/rtpose_b200.h must be consumable by a C compiler (the FFI boundary, no C++ or torch types) and a C program
    linked against librtpose_b200.so must be able to call it: version, error string, and the host-only .npy probe."""
    import shutil
    import subprocess
    import numpy as np
    from rtpose_b200 import lib
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists(lib.LIB_PATH):
        pytest.skip("needs gcc and the built library")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "bind.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "rtpose_b200.h"
int main(int argc, char** argv) {
  rtp_npy_info info;
  rtp_conv_desc d;
  memset(&d, 0, sizeof d);
  if (argc < 2) return 9;
  if (rtp_version() <= 0) return 10;
  if (rtp_npy_probe("/nonexistent/cube.npy", &info) >= 0) return 11;
  if (strstr(rtp_last_error(), "cannot open") == NULL) return 12;
  if (rtp_npy_probe(argv[1], &info) != 0) return 13;
  printf("%d %lld %lld %lld %s %d %d %d %d %d %d %d %d %d\n", info.ndim, (long long)info.shape[0], (long long)info.shape[1],
         (long long)info.shape[2], info.descr, (int)sizeof(rtp_p8), (int)sizeof(rtp_conv_desc), (int)sizeof(rtp_npy_info),
         (int)sizeof(rtp_conv_k3s1_desc), (int)sizeof(rtp_wgrad_desc), (int)sizeof(rtp_fuse_desc), (int)sizeof(rtp_conat_desc),
         (int)sizeof(rtp_pack_job), (int)offsetof(rtp_pack_job, block0));
  printf("%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", (int)offsetof(rtp_conv_desc, w), (int)offsetof(rtp_conv_desc, Cin),
         (int)offsetof(rtp_conv_desc, ntaps), (int)offsetof(rtp_conv_desc, tz), (int)offsetof(rtp_conv_desc, wt),
         (int)offsetof(rtp_conv_desc, RZ), (int)offsetof(rtp_conv_desc, accumulate), (int)offsetof(rtp_conv_k3s1_desc, stat_mode),
         (int)offsetof(rtp_conv_k3s1_desc, stat_ws), (int)offsetof(rtp_conv_k3s1_desc, tap_mask), (int)offsetof(rtp_conv_k3s1_desc, debug),
         (int)offsetof(rtp_wgrad_desc, tc), (int)offsetof(rtp_wgrad_desc, workspace), (int)offsetof(rtp_fuse_desc, low),
         (int)offsetof(rtp_fuse_desc, bias), (int)offsetof(rtp_npy_info, data_offset), (int)offsetof(rtp_conat_desc, c_low),
         (int)offsetof(rtp_conat_desc, K));
  return 0;
}
''')
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"),
                    str(src)], check=True)
    exe = tmp_path / "bind"
    libdir = os.path.dirname(lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir, "-lrtpose_b200",
                    "-Wl,-rpath," + libdir], check=True)
    cube = tmp_path / "c.npy"
    np.save(cube, np.zeros((4, 6, 8), dtype=np.float16))
    r = subprocess.run([str(exe), str(cube)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    lines = r.stdout.strip().splitlines()
    out = lines[0].split()
    assert out[:5] == ["3", "4", "6", "8", "<f2"]
    offs = [int(v) for v in lines[1].split()]
    L = lib
    assert offs == [L.ConvDesc.w.offset, L.ConvDesc.Cin.offset, L.ConvDesc.ntaps.offset, L.ConvDesc.tz.offset, L.ConvDesc.wt.offset,
                    L.ConvDesc.RZ.offset, L.ConvDesc.accumulate.offset, L.ConvK3S1Desc.stat_mode.offset, L.ConvK3S1Desc.stat_ws.offset,
                    L.ConvK3S1Desc.tap_mask.offset, L.ConvK3S1Desc.debug.offset, L.WgradDesc.tc.offset, L.WgradDesc.workspace.offset,
                    L.FuseDesc.low.offset, L.FuseDesc.bias.offset, L.NpyInfo.data_offset.offset, L.ConatDesc.c_low.offset,
                    L.ConatDesc.K.offset]
    # the ctypes mirrors in rtpose_b200/lib.py have the C compiler's struct sizes (field order, padding)
    import ctypes as C
    assert [int(v) for v in out[5:]] == [C.sizeof(lib.P8Struct), C.sizeof(lib.ConvDesc), C.sizeof(lib.NpyInfo),
                                         C.sizeof(lib.ConvK3S1Desc), C.sizeof(lib.WgradDesc), C.sizeof(lib.FuseDesc), C.sizeof(lib.ConatDesc),
                                         C.sizeof(lib.PackJob), lib.PackJob.block0.offset]


@pytest.mark.skipif(not os.path.isdir("/root/reference/det3d"), reason="reference tree not present (GPU box)")
def test_install_as_det3d_does_not_shadow_the_reference_package():
    """INTEGRATION.md §1 calls install_as_det3d() at the top of tools/train.py, BEFORE det3d is imported: the reference's
    own packages (det3d.utils, det3d.torchie, det3d.datasets, ...) must stay reachable afterwards.  This container lacks
    some of the reference's third-party dependencies (terminaltables, spconv), so `import det3d.torchie` may still fail —
    but only on a third-party name, never because a det3d.* module was replaced by an empty shell."""
    import subprocess
    code = r'''
import sys
sys.path.insert(0, "/root/reference"); sys.path.insert(0, %r)
import rtpose_b200.det3d_compat as b
d = b.install_as_det3d()
assert d.__file__.startswith("/root/reference/det3d"), d
from det3d.models import build_detector
assert build_detector is b.build_detector
import det3d.utils
assert any(p.startswith("/root/reference/det3d/utils") for p in det3d.utils.__path__), det3d.utils.__path__
assert hasattr(det3d.utils, "Registry") and hasattr(det3d.utils, "build_from_cfg")
import det3d.utils.dist                      # a sub-package of the reference that needs no third-party module
for name in ("det3d.torchie", "det3d.datasets", "det3d.torchie.utils"):
    try:
        __import__(name)
    except ModuleNotFoundError as ex:
        assert not (ex.name or "").startswith("det3d"), (name, ex.name)   # only a missing third-party dependency
from det3d.ops.dcn import DeformConv
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
