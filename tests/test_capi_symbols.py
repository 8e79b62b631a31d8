"""The C-ABI library loads without a GPU and exports every symbol include/photic_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np

from conftest import ROOT, bits_equal


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "photic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(product_lib):
    from photic_b200 import capi
    names = _declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(product_lib, n), f"{n} declared in photic_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_no_cpu_fallback_without_device(product_lib):
    """Without a usable device the context cannot be created; nothing silently runs on the CPU."""
    import torch
    from photic_b200 import capi
    if torch.cuda.is_available():
        return
    ctx = C.c_void_p()
    assert product_lib.phb_ctx_create(0, C.byref(ctx)) == 2  # PHB_ENODEVICE
    assert b"no CPU fallback" in product_lib.phb_error_string(2)
    assert product_lib.phb_device_count() == 0


def test_descriptor_validation(product_lib):
    from photic_b200 import capi, scene
    d = capi.desc_from_spec(scene.CONFIGS["murion"])
    out = np.zeros((16, 8, 12))
    aux = np.zeros(33)
    assert product_lib.phb_band_tables(C.byref(d), out.ctypes.data_as(capi._dp), aux.ctypes.data_as(capi._dp)) == 0
    d.n_scenes = 17
    assert product_lib.phb_band_tables(C.byref(d), out.ctypes.data_as(capi._dp), aux.ctypes.data_as(capi._dp)) == 1
    assert product_lib.phb_debug_record_len(C.byref(d)) == 0


def test_host_band_tables_equal_oracle(product_lib, oracle_port):
    """Host-only part of the product (scene constants, samodel.c:505-618) against the oracle."""
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    for name in ("murion", "exmouth", "pilbara"):
        spec = scene.CONFIGS[name]
        d = capi.desc_from_spec(spec)
        out = np.zeros((16, 8, 12))
        aux = np.zeros(33)
        assert product_lib.phb_band_tables(C.byref(d), out.ctypes.data_as(capi._dp), aux.ctypes.data_as(capi._dp)) == 0
        t0, t1 = oracle_port.tables(SceneCfg.from_spec(spec))
        ns = spec.n_dates
        assert bits_equal(out[:ns, :4, :7], t0).all()
        assert bits_equal(aux[:1 + 2 * ns], t1).all()


def test_struct_layout_matches_header(product_lib):
    """ctypes mirrors of the POD structs have the sizes the C compiler gives them."""
    import subprocess, tempfile
    from photic_b200 import capi
    code = '#include <stdio.h>\n#include "photic_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(phb_scene_desc), sizeof(phb_outputs), sizeof(phb_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.SceneDesc), C.sizeof(capi.Outputs), C.sizeof(capi.Stats)]
