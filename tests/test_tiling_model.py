"""Host-side model of the kernels' tile -> pixel bookkeeping (conv_gemm.cu decode_tile / epilogue_tile, wgrad.cu chunk
coordinates): every output pixel of a ragged batch is produced exactly once, CTA pairs split the pixel block 4 + 4
images, and every pixel row is reduced exactly once by the weight-gradient chunks.  Pure Python - a regression guard
for the index arithmetic that the GPU tests exercise end to end."""
import itertools

import pytest

IMG_PIX, BN, IMG_PER_TILE, ROWS = 130, 208, 8, 26


def conv_tile_rows(nt, n_img):
    """rows written by the epilogue of pixel tile nt (conv mode)."""
    out = []
    for nl in range(IMG_PER_TILE):
        img = (nt // 5) * IMG_PER_TILE + nl
        if img < n_img:
            row0 = img * IMG_PIX + (nt % 5) * ROWS
            out += list(range(row0, row0 + ROWS))
    return out


@pytest.mark.parametrize("n_img", [1, 7, 8, 9, 21, 64, 2240 // 10])
def test_conv_tiles_cover_every_pixel_once(n_img):
    n_tiles = ((n_img + IMG_PER_TILE - 1) // IMG_PER_TILE) * 5
    rows = list(itertools.chain.from_iterable(conv_tile_rows(nt, n_img) for nt in range(n_tiles)))
    assert sorted(rows) == list(range(n_img * IMG_PIX))


def test_tmem_column_to_pixel_mapping_matches_tma_box_order():
    """TMA box {32 ch, 13 w, 2 h, 8 img} lands in smem as row = img*26 + h*13 + w = the TMEM column; the epilogue
    maps column c -> image c // 26, pixel (h0 + (c % 26) // 13, (c % 26) % 13)."""
    h0 = 4
    for c in range(BN):
        img, r = divmod(c, ROWS)
        hh, w = divmod(r, 13)
        assert c == img * 26 + hh * 13 + w
        assert (h0 + hh) * 13 + w == h0 * 13 + r          # row0 + r in the kernel


@pytest.mark.parametrize("m_tiles,n_tiles,splits,ncta", [(8, 10, 1, 2), (8, 10, 1, 1), (1, 11, 13, 1), (4, 3, 1, 2)])
def test_work_items_enumerate_every_tile_once(m_tiles, n_tiles, splits, ncta):
    mg = m_tiles // ncta
    seen = set()
    for item in range(mg * n_tiles * splits):
        for rank in range(ncta):
            mt = (item % mg) * ncta + rank
            r = item // mg
            seen.add((mt, r % n_tiles, r // n_tiles))
    assert len(seen) == m_tiles * n_tiles * splits


def test_cta_pair_splits_the_pixel_block_by_images():
    nt = 7
    for rank in (0, 1):
        first_img = (nt // 5) * IMG_PER_TILE + rank * (IMG_PER_TILE // 2)
        assert first_img == 8 + 4 * rank          # CTA r stages images n0+4r .. n0+4r+3 = B rows 104r .. 104r+103


@pytest.mark.parametrize("n_img", [4, 5, 11, 64])
def test_wgrad_chunks_reduce_every_pixel_row_once(n_img):
    n_chunks = ((n_img + 3) // 4) * 5
    rows = []
    for c in range(n_chunks):
        h0, n0 = (c % 5) * 2, (c // 5) * 4
        for n in range(n0, n0 + 4):
            if n < n_img:                          # out-of-range images are TMA zero fill
                rows += [n * IMG_PIX + (h0 + hh) * 13 + w for hh in range(2) for w in range(13)]
    assert sorted(rows) == list(range(n_img * IMG_PIX))


def test_wgrad_split_plan_owns_every_chunk():
    """host logic of vnqa_conv_wgrad: every split owns >= 1 chunk (an empty split never signals its accumulator)."""
    for n_chunks, base_items, groups in [(10, 1, 148), (2800, 576, 148), (2800, 144, 74), (22, 1040, 148), (5, 9, 148)]:
        chunk_bytes = 104 * 4 * 2048
        cps = int(48.0e6 / chunk_bytes)
        cps = min(cps, n_chunks * base_items // (2 * groups))
        cps = max(cps, 4)
        cps = min(cps, n_chunks)
        splits = (n_chunks + cps - 1) // cps
        assert all(s * cps < n_chunks for s in range(splits))
        assert splits * cps >= n_chunks
