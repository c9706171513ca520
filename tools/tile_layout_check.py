"""Host-side model of the shared-memory tile layouts the pair GEMM's splitter warps read and write (gemm_tcgen05_2cta.cu):
the TMA swizzles of the raw f32 tiles, the two K-major 16-bit tile layouts, and the canonical UMMA K-major read order of
cute/atom/mma_traits_sm100.hpp.  It checks that
  * the splitter's read addresses find element (row, k) in a raw tile written by TMA with SWIZZLE_128B (K-major operand) or
    SWIZZLE_128B_ATOM_32B (MN-major operand, 4 boxes of [32 k][32 rows]);
  * a 16-bit tile written with bf_chunk_offset() is read back in (row, k) order by a SWIZZLE_64B descriptor (SBO 512, K slice
    i at +32 i bytes) or an unswizzled one (SBO 128, LBO 2048, slice i at +4096 i);
  * every warp-wide access is bank-conflict free (8 distinct 16-byte slots per quarter warp for 128-bit accesses, 32 distinct
    words for 32-bit ones);
  * the staged epilogue's writes (row `lane`, 16-byte chunk j in slot j ^ (lane & 7)) are what a SWIZZLE_128B tensor store reads.
Pure Python: run it anywhere.  Used by tests/test_tile_layouts.py."""

sw128 = lambda a: a ^ (((a >> 7) & 7) << 4)       # Swizzle<3,4,3> on byte addresses
sw128_32 = lambda a: a ^ (((a >> 7) & 3) << 5)    # Swizzle<2,5,2>: 128B swizzle with 32-byte atoms
sw64 = lambda a: a ^ (((a >> 7) & 3) << 4)        # Swizzle<2,4,3>

ROWS, BK = 128, 32


def raw_k_major():
    """byte address -> (row, k) of the f32 element TMA puts there (box: 32 k contiguous x 128 rows, SWIZZLE_128B)"""
    return {sw128(r * 128 + k * 4): (r, k) for r in range(ROWS) for k in range(BK)}


def raw_mn_major():
    """4 boxes of [32 k rows][32 rows of the operand contiguous], SWIZZLE_128B_ATOM_32B"""
    return {(r >> 5) * 4096 + sw128_32(k * 128 + (r & 31) * 4): (r, k) for r in range(ROWS) for k in range(BK)}


def splitter_reads(mn_major, row):
    """(address, k) pairs split_row_bf16 issues for one row; 128-bit reads are listed per 4-byte element"""
    out = []
    if mn_major:
        base, c = (row >> 5) * 4096 + (row & 7) * 4, (row & 31) >> 3
        for k in range(BK):
            out.append((base + k * 128 + ((c ^ (k & 3)) << 5), k))
    else:
        for c in range(8):
            a = row * 128 + ((c ^ (row & 7)) << 4)
            out += [(a + 4 * e, 4 * c + e) for e in range(4)]
    return out


def chunk_offset(swizzled, row, j):
    return row * 64 + ((j ^ ((row >> 1) & 3)) << 4) if swizzled else j * 2048 + row * 16


def umma_read_address(swizzled, i, row, kk):
    """byte address of element (row, 16 i + kk) as the canonical K-major layouts define it"""
    if swizzled:
        return sw64(i * 32 + (row % 8) * 64 + (row // 8) * 512 + (kk // 8) * 16 + (kk % 8) * 2)
    return i * 4096 + (row % 8) * 16 + (row // 8) * 128 + (kk // 8) * 2048 + (kk % 8) * 2


def check():
    problems = []
    for mn, table in ((False, raw_k_major()), (True, raw_mn_major())):
        for row in range(ROWS):
            for a, k in splitter_reads(mn, row):
                if table.get(a) != (row, k):
                    problems.append(f"raw tile mn_major={mn}: row {row} k {k} read at {a} finds {table.get(a)}")
    for swz in (True, False):
        tile = {}
        for row in range(ROWS):
            for j in range(4):
                for e in range(8):
                    tile[chunk_offset(swz, row, j) + 2 * e] = (row, 8 * j + e)
        if len(tile) != ROWS * BK:
            problems.append(f"16-bit tile swizzled={swz}: overlapping writes")
        for i in range(2):
            for row in range(ROWS):
                for kk in range(16):
                    if tile.get(umma_read_address(swz, i, row, kk)) != (row, 16 * i + kk):
                        problems.append(f"16-bit tile swizzled={swz}: slice {i} row {row} k {kk}")
    # bank conflicts
    slots = lambda addrs: len({(a % 128) // 16 for a in addrs})
    for q in range(0, ROWS, 8):
        for c in range(8):
            if slots([r * 128 + ((c ^ (r & 7)) << 4) for r in range(q, q + 8)]) != 8:
                problems.append("K-major raw read conflicts")
        for j in range(4):
            for swz in (True, False):
                if slots([chunk_offset(swz, r, j) for r in range(q, q + 8)]) != 8:
                    problems.append(f"16-bit tile write conflicts swizzled={swz}")
        if q % 32 == 0:  # epilogue staging: lanes 0..31 = rows of a 32 x 32 piece
            pass
    for w in range(0, ROWS, 32):
        for k in range(BK):
            words = {(((r >> 5) * 4096 + (r & 7) * 4 + k * 128 + ((((r & 31) >> 3) ^ (k & 3)) << 5)) % 128) // 4 for r in range(w, w + 32)}
            if len(words) != 32:
                problems.append("MN-major raw read conflicts")
    # staged epilogue: thread `lane` writes chunk j of its row to lane*128 + ((j ^ (lane & 7)) << 4); the TMA reads SWIZZLE_128B
    for lane in range(32):
        for j in range(8):
            if lane * 128 + ((j ^ (lane & 7)) << 4) != sw128(lane * 128 + j * 16):
                problems.append("epilogue staging does not match SWIZZLE_128B")
    for q in range(0, 32, 8):
        for j in range(8):
            if slots([l * 128 + ((j ^ (l & 7)) << 4) for l in range(q, q + 8)]) != 8:
                problems.append("epilogue staging write conflicts")
    return problems


if __name__ == "__main__":
    p = check()
    print("tile layouts: OK" if not p else "\n".join(p[:20]))
    raise SystemExit(1 if p else 0)
