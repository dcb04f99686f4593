"""Desk check (CPU, no extension call) of the lane -> (tile row, 16-byte unit) mapping of the phase-aligned fetch
scheme of vision3d_b200/csrc/sparse_conv_tc.cu (kVer >= 2): the formulas below are the kernel's, restated; the test
proves that every 16-byte unit of both A tiles (h1, h2) of a K = 64 slot is written exactly once, at the 128-byte
swizzled position the UMMA descriptor reads, from the rule entry of the right tile row and the right source bytes."""
import pytest

K_TILE_M = 128
K_A_TILE = K_TILE_M * 128


def staged_position(row):
    """where the rule entry of tile row `row` is staged inside a 128-entry rule row (kernel: `spos`)"""
    return ((row & 7) * 2 + (row >> 6)) * 8 + ((row >> 3) & 7)


@pytest.mark.parametrize("cin", [16, 32, 64])
def test_phase_aligned_fetch_covers_the_a_tiles_exactly_once(cin):
    assert sorted(staged_position(r) for r in range(K_TILE_M)) == list(range(K_TILE_M))   # a permutation
    row_at = {staged_position(r): r for r in range(K_TILE_M)}
    written = {}
    for fw in range(8):                      # fetch warp = swizzle phase of every row it copies
        for lane in range(32):
            half, part, unit = lane >> 4, (lane >> 3) & 1, lane & 7
            off = (unit * 8) // cin          # which of the 64/cin stacked kernel offsets this unit belongs to
            src_byte = part * (2 * cin) + ((unit * 8) % cin) * 2
            dst_lane = part * K_A_TILE + (64 * half + fw) * 128 + ((unit ^ fw) << 4)
            idx_pos = (fw * 2 + half) * 8
            for j in range(8):
                dst = dst_lane + j * 1024    # the LDGSTS immediate
                row = row_at[idx_pos + j]    # the rule entry the lane uses for this copy
                assert dst not in written
                written[dst] = (row, off, src_byte)
                tile, r, u = dst // K_A_TILE, (dst % K_A_TILE) // 128, ((dst % 128) >> 4) ^ (((dst % K_A_TILE) // 128) & 7)
                assert (tile, r, u) == (part, row, unit)
                # unit u of a tile row holds K elements [8u, 8u + 8) = channels (8u % cin).. of stacked offset 8u // cin,
                # i.e. bytes [2 * (8u % cin), +16) of the h1 half (part 0) or the h2 half (part 1) of the packed source row
                assert off == (8 * u) // cin and src_byte == part * 2 * cin + 2 * ((8 * u) % cin)
    assert len(written) == 2 * K_TILE_M * 8


def test_conv_variant_selection_from_environment():
    """v3d_sparse_conv_tc_variant (no device needed): fetch scheme + 8 * cg + 16 * spin; unknown values fall back to the
    built-in defaults (scheme 4 + cp.async.cg, the variant measured fastest: profiles/r02x_conv_variants.txt)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from vision3d_b200 import _lib; "
            "print(_lib.load().v3d_sparse_conv_tc_variant())" % root)

    def variant(**env):
        e = {k: v for k, v in os.environ.items() if not k.startswith("V3D_TC_")}
        e.update(env)
        return int(subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True,
                                  check=True).stdout.strip())

    assert variant() == 4 + 8
    assert variant(V3D_TC_FETCH="1", V3D_TC_CG="0") == 1
    assert variant(V3D_TC_FETCH="2", V3D_TC_WAIT="1") == 2 + 8 + 16
    assert variant(V3D_TC_FETCH="7", V3D_TC_CG="x") == 4 + 8


def test_bench_reports_the_conv_variant():
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    v = bench._conv_variant()
    assert set(v) >= {"fetch_scheme", "cp_async_cg", "spin_wait"} and v["fetch_scheme"] in (1, 2, 4, 5)
