"""The C-ABI library loads on a CPU-only box and exports every symbol include/tsdf_b200.h declares.
No compute calls: without a GPU every compute entry point must fail loudly, never fall back."""
import ctypes
import math
import os
import re

import pytest

import tracking_sdf_b200 as T
from tracking_sdf_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tsdf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsdf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    L = T.load_library()
    names = declared_symbols()
    assert len(names) >= 45
    bound = {n for n, _, _ in capi.PROTOTYPES}
    for n in names:
        assert hasattr(L, n), "library does not export " + n
        assert n in bound, "capi.py does not bind " + n
    assert L.tsdf_abi_version() == 1


def test_default_config_is_the_reference_node_setup():
    # sdf_reconstruction.cpp:83-88
    c = T.default_config()
    assert (c.m, c.width, c.height, c.depth) == (256, 6.0, 6.0, 3.5)
    assert tuple(c.origin) == (-3.0, -3.0, -0.5)
    assert c.distance_delta == pytest.approx(0.3) and c.distance_epsilon == pytest.approx(0.025)
    assert c.gauss_newton_max_iteration == 20 and c.maximum_twist_diff == pytest.approx(0.001)
    assert c.v_h == 1.0 and c.w_h == pytest.approx(0.01) and c.pixel_stride == 3
    assert c.metric == T.POINT_TO_PLANE and (c.image_width, c.image_height) == (640, 480)


def test_no_cpu_fallback():
    L = T.load_library()
    if L.tsdf_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(T.TsdfError) as e:
        T.Tsdf(m=64)
    assert e.value.status == 3 and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_rejected_without_a_device():
    L = T.load_library()
    h = ctypes.c_void_p()
    for kw in [dict(m=30), dict(m=4), dict(pixel_stride=0), dict(metric=7), dict(n_shards=0), dict(n_shards=2, shard_rank=2)]:
        c = T.default_config(**kw)
        st = L.tsdf_create(ctypes.byref(c), ctypes.byref(h))
        assert st == 1 and not h.value, kw
    assert L.tsdf_create(None, ctypes.byref(h)) == 1
    assert L.tsdf_set_pose(None, None, None) == 1
    assert L.tsdf_destroy(None) == 0


def test_slab_plan_partitions_and_halo():
    # SURVEY.md §8e + hard part 5: contiguous z-slabs; halo covers the stencil incl. the rotational reach
    for m, G in [(512, 1), (1024, 2), (1024, 8), (2048, 8), (256, 3)]:
        own = []
        for r in range(G):
            p = capi.slab_plan(T.default_config(m=m, n_shards=G, shard_rank=r))
            own.append(p["own"])
            assert p["stored"][0] == max(0, p["own"][0] - p["halo"]) and p["stored"][1] == min(m, p["own"][1] + p["halo"])
            if G == 1:
                assert p["halo"] == 0 and p["stored"] == (0, m)
            else:
                assert p["halo"] >= 2 + math.ceil(0.01 * 6.0 / (3.5 / m))     # w_h * max(width, height) along z
        assert own[0][0] == 0 and own[-1][1] == m
        assert all(own[r][1] == own[r + 1][0] for r in range(G - 1))
    p = capi.slab_plan(T.default_config(m=1024, n_shards=4, shard_rank=1, halo=5))
    assert p["halo"] == 5 and p["stored"] == (256 - 5, 512 + 5)
