"""The C-ABI library loads without a GPU and exports exactly what include/ngw.h declares; the Python mirrors of the
header's enums / struct layout agree with the C compiler's view."""
import ctypes as C
import os
import re

import pytest

from gym_novel_gridworlds_b200 import capi, opcodes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, 'include', 'ngw.h')).read()


def _declared_functions():
    body = HEADER[HEADER.index('typedef struct ngw_handle ngw_handle;'):]
    return set(re.findall(r'^(?:int|void|int64_t|const char\*)\s+(ngw_[a-z_]+)\(', body, re.M))


def test_header_and_binding_declare_the_same_entry_points():
    assert _declared_functions() == set(capi.EXPORTS)


@pytest.mark.skipif(not os.path.exists(capi.LIB_PATH), reason="libngw_b200.so not built (run __graft_entry__.build())")
def test_library_loads_and_exports_every_symbol():
    lib = capi.load_library()
    for name in _declared_functions():
        assert getattr(lib, name) is not None
    assert lib.ngw_abi_version() == capi.ABI_VERSION == int(re.search(r'#define NGW_ABI_VERSION (\d+)', HEADER).group(1))
    assert lib.ngw_last_error() is not None


def test_missing_library_fails_loudly():
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        capi.load_library('/nonexistent/libngw_b200.so')


def test_python_constants_match_header():
    for name in ('MAX_ITEMS', 'MAX_ACTIONS', 'MAX_RECIPES', 'MAX_RECIPE_INPUTS', 'MAX_LAYERS', 'MAX_PLACE',
                 'MAX_RESET_OPS', 'MAX_MAP_SIZE'):
        assert getattr(opcodes, name) == int(re.search(r'#define NGW_%s (\d+)' % name, HEADER).group(1))

    def enum_values(enum_name):
        body = re.search(r'enum %s \{(.*?)\};' % enum_name, HEADER, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        return {m.group(1): int(m.group(2)) for m in re.finditer(r'NGW_([A-Z_]+) = (\d+)', body)}

    for k, v in enum_values('ngw_op').items():
        assert getattr(opcodes, k) == v
    for k, v in enum_values('ngw_break_variant').items():
        assert getattr(opcodes, k) == v
    for k, v in enum_values('ngw_layer').items():
        assert getattr(opcodes, k) == v
    for k, v in enum_values('ngw_reset_kind').items():
        assert getattr(opcodes, k) == v


def test_ctypes_config_layout_matches_the_c_compiler():
    from oracle import oracle_lib
    assert oracle_lib.lib().ngo_sizeof_config() == C.sizeof(capi.ConfigC)


@pytest.mark.skipif(not os.path.exists(capi.LIB_PATH), reason="libngw_b200.so not built")
def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import bench
    from gym_novel_gridworlds_b200.compiler import compile_chain
    from gym_novel_gridworlds_b200.runtime import BatchHandle
    cc = compile_chain(bench.build_c2_chain())
    with pytest.raises(RuntimeError, match="no CPU path"):
        BatchHandle([cc], 64)
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.ngw_create(C.byref(h), C.byref(cc.c), 1, 64, 10, 0, 0, 0) != 0
    assert lib.ngw_last_error() and not h.value
    env = bench.build_c2_chain()
    with pytest.raises(RuntimeError):
        env.reset()


@pytest.mark.skipif(not os.path.exists(capi.LIB_PATH), reason="libngw_b200.so not built")
def test_lidar_path_selection_needs_no_gpu():
    """ngw_lidar_path: the reference's 8-beam geometry takes the line-gather path (3) on every grid size; other beam
    counts walk the generic LUT (1); no LidarInFront wrapper -> 0."""
    import scenarios
    from gym_novel_gridworlds_b200.compiler import compile_chain
    lib = capi.load_library()
    ns = scenarios.b200_namespace()
    for ms in (9, 10, 11, 17, 32, 33, 40, 64):
        cc = compile_chain(scenarios.build_chain(ns, {'env': scenarios.POGO, 'map_size': ms, 'chain': [['lidar', 8]]}))
        assert lib.ngw_lidar_path(C.byref(cc.c), ms) == 3, ms
    for beams in (1, 5, 16):
        cc = compile_chain(scenarios.build_chain(ns, {'env': scenarios.BOW, 'map_size': 12, 'chain': [['lidar', beams]]}))
        assert lib.ngw_lidar_path(C.byref(cc.c), 12) == 1, beams
    cc = compile_chain(scenarios.build_chain(ns, {'env': scenarios.POGO, 'map_size': 10, 'chain': []}))
    assert lib.ngw_lidar_path(C.byref(cc.c), 10) == 0
