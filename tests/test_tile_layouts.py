"""Host model of the pair GEMM's shared-memory tile layouts (tools/tile_layout_check.py): swizzle decode of the raw TMA tiles,
the K-major 16-bit tiles against the canonical UMMA read order, the epilogue staging against SWIZZLE_128B, bank conflicts.
CPU only."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_splitter_and_epilogue_address_algebra():
    spec = importlib.util.spec_from_file_location("tile_layout_check", os.path.join(ROOT, "tools", "tile_layout_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check() == []


def test_conv_tile_index_algebra():
    """Host model of the register-tiled Conv2D kernels (tools/conv_tile_check.py): the bordered planar gradient tile, the 4 x 4 dx
    tiles and their windows, the (channel, filter row) x output-row dw mapping and the forward strips reproduce the oracle's literal
    lowering; no access leaves its buffer; the pitches keep a quarter warp's 16-byte accesses in 8 distinct bank groups."""
    spec = importlib.util.spec_from_file_location("conv_tile_check", os.path.join(ROOT, "tools", "conv_tile_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check() == []


def test_activation_warps_decomposition_and_handover():
    """Host model of the pair GEMM's activation warps (tools/act_warps_check.py): every 16-byte vector of a CTA's part of a tile is
    visited exactly once on both index paths, fast-path accesses are 512 contiguous bytes, and the z_done / act_done mbarrier
    hand-over neither deadlocks nor lets a barrier run two phases ahead of a waiter under randomised interleavings."""
    spec = importlib.util.spec_from_file_location("act_warps_check", os.path.join(ROOT, "tools", "act_warps_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check() == []
