"""The wire / on-disk container around compressed streams (include/wah_b200.h, host only: runs without a GPU)."""
import ctypes

import numpy as np
import pytest

import datagen
import gpu_wah_b200 as wah
import oracle_lib as orc


def _streams(mode):
    cols = [datagen.uniform(992 * 3 + 17, 0.001, 1), np.zeros(5000, dtype=np.uint32), datagen.clustered(4000, 0.3, 200, 2),
            datagen.uniform(64, 0.5, 3)]
    comp = [orc.compress(c, mode) for c in cols]
    offs = np.concatenate([[0], np.cumsum([c.size for c in comp])]).astype(np.uint64)
    return cols, np.concatenate(comp), offs


@pytest.mark.parametrize("mode", [wah.WAH_BLOCK1024, wah.WAH_CANONICAL])
def test_round_trip_of_several_streams(mode):
    cols, words, offs = _streams(mode)
    buf = wah.container_pack(words, mode, 0, offs)
    assert buf.size == wah.lib.wah_container_bytes(len(cols), words.size) and bytes(buf[:8]) == b"WAHB200\0"
    m, wps, o, w = wah.container_unpack(buf.tobytes())
    assert m == mode and wps == 0
    assert np.array_equal(o, offs) and np.array_equal(w, words)
    for j, c in enumerate(cols):   # every stream decodes on its own
        assert np.array_equal(orc.decompress(w[int(o[j]):int(o[j + 1])])[:c.size], c)


def test_single_and_empty_streams():
    data = datagen.uniform(2048, 0.01, 9)
    cw = orc.compress(data, 0)
    m, wps, o, w = wah.container_unpack(wah.container_pack(cw, 0, data.size))
    assert (m, wps, o.tolist()) == (0, data.size, [0, cw.size]) and np.array_equal(w, cw)
    m, wps, o, w = wah.container_unpack(wah.container_pack(np.zeros(0, dtype=np.uint32), 1, 0))
    assert (m, wps, o.tolist(), w.size) == (1, 0, [0, 0], 0)


def test_damage_is_detected():
    _, words, offs = _streams(0)
    good = wah.container_pack(words, 0, 0, offs).copy()
    hb = good.size - 4 * words.size

    def code(buf):
        with pytest.raises(wah.WahError) as e:
            wah.container_unpack(buf)
        return e.value.code

    bad = good.copy(); bad[3] ^= 1                      # magic
    assert code(bad) == 5
    bad = good.copy(); bad[8] = 9                       # version
    assert code(bad) == 5
    bad = good.copy(); bad[12] = 7                      # mode
    assert code(bad) == 5
    bad = good.copy(); bad[hb + 40] ^= 0x10             # one payload bit
    assert code(bad) == 5
    bad = good.copy(); bad[hb + 4: hb + 8], bad[hb + 8: hb + 12] = good[hb + 8: hb + 12].copy(), good[hb + 4: hb + 8].copy()
    assert code(bad) == 5                               # two words swapped: the checksum is position dependent
    bad = good.copy(); bad[64 + 8] += 1                 # offset table
    assert code(bad) == 5
    assert code(good[:-4].copy()) == 5                  # truncated payload
    assert code(good[:40].copy()) == 5                  # truncated header
    assert wah.container_unpack(good)[0] == 0           # and the untouched one still opens


def test_pack_rejects_bad_arguments():
    w = np.arange(10, dtype=np.uint32)
    with pytest.raises(wah.WahError):
        wah.container_pack(w, 0, 0, [0, 4, 3, 10])      # offsets decrease
    with pytest.raises(wah.WahError):
        wah.container_pack(w, 0, 0, [1, 10])            # does not start at 0
    with pytest.raises(wah.WahError):
        wah.container_pack(w, 5)                        # unknown mode
    offs = (ctypes.c_uint64 * 2)(0, 10)
    rc = wah.lib.wah_container_pack(ctypes.create_string_buffer(64).raw, 64, 0, 1, 0, offs, w.ctypes.data)
    assert rc == 4                                      # WAH_ERR_CAPACITY
