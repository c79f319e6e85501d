"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/ovn_b200.h
declares; the product path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def built():
  from overlapnet_b200 import build
  return build.build()


def header_functions():
  src = open(os.path.join(ROOT, 'include', 'ovn_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(ovn_[a-z0-9_A-Z]+)\s*\(', src)))


def test_header_and_binding_agree(built):
  from overlapnet_b200 import _cabi
  assert sorted(_cabi.SYMBOLS) == header_functions()


def test_library_exports_every_declared_symbol(built):
  lib = C.CDLL(built)
  for name in header_functions():
    assert hasattr(lib, name), name
  from overlapnet_b200 import _cabi
  assert _cabi.lib().ovn_abi_version() == _cabi.OVN_ABI_VERSION


def test_config_struct_layout_matches_header(built):
  from overlapnet_b200 import _cabi
  cfg = _cabi.OvnConfig()
  _cabi.lib().ovn_default_config(C.byref(cfg))
  assert (cfg.abi_version, cfg.proj_H, cfg.proj_W) == (1, 64, 900)
  assert (cfg.fov_up_deg, cfg.fov_down_deg, cfg.max_range) == (3.0, -25.0, 50.0)
  assert (cfg.use_depth, cfg.use_normals, cfg.n_prob_channels, cfg.use_intensity) == (1, 1, 0, 0)
  assert list(cfg.strides_layer1) == [2, 2] and cfg.additional_unsymmetric_layer3a == 1
  assert (cfg.leg_output_width, cfg.conv1size, cfg.precision) == (360, 15, 1)
  assert (cfg.max_batch_scans, cfg.max_batch_pairs) == (16, 1101)
  assert C.sizeof(_cabi.OvnConfig) == 18 * 4


def test_no_cpu_fallback(built):
  """Without a CUDA device ovn_create must fail with OVN_ERR_NO_DEVICE, never compute on the CPU."""
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  from overlapnet_b200 import _cabi
  L = _cabi.lib()
  cfg = _cabi.OvnConfig()
  L.ovn_default_config(C.byref(cfg))
  h = C.c_void_p(0)
  st = L.ovn_create(C.byref(cfg), C.byref(h))
  assert st == -5 and L.ovn_status_string(st) == b'OVN_ERR_NO_DEVICE'
  assert b'no CPU fallback' in L.ovn_last_error(None)
  from overlapnet_b200.engine import Engine
  with pytest.raises(_cabi.OvnError):
    Engine()


def test_product_never_imports_oracle():
  """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
  pkg = os.path.join(ROOT, 'overlapnet_b200')
  for dp, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        txt = open(os.path.join(dp, f)).read()
        assert 'import oracle' not in txt and 'from oracle' not in txt, os.path.join(dp, f)
