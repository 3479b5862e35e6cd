"""The C-ABI library loads on a machine without a GPU, exports every symbol include/gpdoemd_b200.h declares,
fails loudly instead of falling back, and its layout helpers describe a consistent DMMA operand packing."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'gpdoemd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gpd_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound():
    from gpdoemd_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), 'declared in the header but not exported: ' + n
    assert set(names) == set(_lib.exported_symbols()), set(names) ^ set(_lib.exported_symbols())


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from gpdoemd_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.gpd_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert len(lib.gpd_last_error()) > 0
    import gpdoemd_b200 as g
    with pytest.raises(_lib.GpdError):
        g.HR(np.zeros((3, 2)), np.ones((3, 2)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gpdoemd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
                assert 'oracle/' not in txt, f


def test_dmma_operand_layouts_are_bijections():
    from gpdoemd_b200 import _lib
    lib = _lib.load()
    # A chunk: 128 rows x 16 k -> 2048 distinct slots; fragment (ms, ks, lane) owns (row, k)
    seen = {}
    for ms in range(16):
        for ks in range(4):
            for lane in range(32):
                idx = lib.gpd_layout_a_index(ms, ks, lane)
                row, col = lib.gpd_layout_a_row(ms, lane), lib.gpd_layout_a_col(ks, lane)
                assert 0 <= idx < 2048 and idx not in seen
                seen[idx] = (row, col)
                assert row == ms * 8 + lane // 4 and col == ks * 4 + lane % 4       # PTX m8n8k4 A fragment
    assert len(seen) == 2048 and len(set(seen.values())) == 2048
    # B chunk: 16 k x 128 columns with the XOR swizzle; a half warp (4 k x 4 t... 16 lanes) hits 16 distinct banks
    idxs = {lib.gpd_layout_b_index(k, t) for k in range(16) for t in range(128)}
    assert idxs == set(range(2048))
    for ks in range(4):
        for t0 in range(0, 128, 8):
            for half in range(2):
                banks = {(lib.gpd_layout_b_index(ks * 4 + lane % 4, t0 + lane // 4) * 2) % 32
                         for lane in range(half * 16, half * 16 + 16)}
                assert len(banks) == 16
    # packed tile runs: forward rows hold k blocks 0..I (diagonal last), backward I..nblk-1 (diagonal first)
    for nblk in (1, 2, 5):
        for bwd in (0, 1):
            offs = []
            for I in range(nblk):
                for kb in range(nblk):
                    o = lib.gpd_layout_tile_offset(nblk, I, kb, bwd)
                    needed = (kb >= I) if bwd else (kb <= I)
                    assert (o >= 0) == needed
                    if o >= 0:
                        offs.append(o)
            assert sorted(offs) == [i * 128 * 128 for i in range(nblk * (nblk + 1) // 2)]
            assert offs == sorted(offs)          # run order == memory order: one sequential stream per item


def test_binary_combo_weights_match_reference_order():
    from gpdoemd_b200 import _lib
    from oracle.routing import sign_table
    lib = _lib.load()
    for nb in (1, 2, 3, 4):
        B = sign_table(nb)
        w = [lib.gpd_binary_combo_weight(nb, v) for v in range(nb)]
        for j, row in enumerate(B):
            assert sum(wv for wv, s in zip(w, row) if s > 0) == j


def test_header_is_valid_c(tmp_path):
    """include/gpdoemd_b200.h is the boundary a C / cgo / JNI host would bind: it must compile as plain C99 and C++17."""
    import subprocess
    src = tmp_path / 'use_header.c'
    src.write_text('#include "gpdoemd_b200.h"\n'
                   'int probe(gpd_ctx* c, gpd_comm* k, gpd_model* const* m, const double* x, double* bv, int64_t* bi) {\n'
                   '    const int32_t kinds[2] = {GPD_HR, GPD_BH};\n'
                   '    char id[GPD_NCCL_ID_BYTES];\n'
                   '    double noise[4] = {0.01, 0, 0, 0.01}, pps[2] = {0.5, 0.5};\n'
                   '    if (gpd_nccl_unique_id(id)) return 1;\n'
                   '    return gpd_score_sharded(c, k, m, 2, x, 0, 100, 1, kinds, 2, noise, pps, 0, 0, bv, bi, 0) == GPD_ERR_SINGULAR;\n'
                   '}\n')
    inc = os.path.join(ROOT, 'include')
    for cmd in (['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', inc, '-c', str(src), '-o', str(tmp_path / 'a.o')],
                ['g++', '-std=c++17', '-Wall', '-Werror', '-x', 'c++', '-I', inc, '-c', str(src), '-o', str(tmp_path / 'b.o')]):
        out = subprocess.run(cmd, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
