"""save_t round trip through the on-disk form (include/annb200_io.h).  Runs without a GPU:
the index comes from the CPU checker, the I/O functions are pure host code of the product
library, and the reloaded index must answer queries exactly like the original."""
import numpy as np
import pytest

from conftest import load_golden, same_bits


@pytest.mark.parametrize("name", ["ragged_f32", "ragged_f64"])
def test_save_roundtrip(tmp_path, oracle_mod, name):
    from approximatenn_b200.api import gpu_backend, save_from_file, save_to_file
    g = load_golden(name)
    orc = oracle_mod.restatement(g["dtype"])
    lib = gpu_backend(g["dtype"])                     # loads without a GPU; only host functions are used
    res = orc.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    path = str(tmp_path / "index.annb")
    save_to_file(lib, res.save, path)
    back = save_from_file(lib, path)
    a, b = res.save, back
    assert (a.tries, a.n, a.k, a.d_short, a.d_long) == (b.tries, b.n, b.k, b.d_short, b.d_long)
    assert np.array_equal(a.par_maxes, b.par_maxes) and np.array_equal(a.graph, b.graph)
    assert same_bits(a.row_means, b.row_means) and same_bits(a.bases, b.bases)
    for t in range(a.tries):
        assert np.array_equal(a.which_par(t), b.which_par(t))
    q = orc.query(b, g["points"], g["y"])             # the reloaded index answers like the reference
    assert np.array_equal(q.ids, g["q_ids"].astype(np.uint64)) and same_bits(q.dists, g["q_dists"])
    back.free(); res.save.free()


def test_wrong_ftype_or_garbage_is_rejected(tmp_path, oracle_mod):
    from approximatenn_b200.api import gpu_backend, save_from_file, save_to_file
    g = load_golden("ragged_f32")
    orc = oracle_mod.restatement(g["dtype"])
    res = orc.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    path = str(tmp_path / "index.annb")
    save_to_file(gpu_backend(np.float32), res.save, path)
    with pytest.raises(OSError):
        save_from_file(gpu_backend(np.float64), path)
    (tmp_path / "junk").write_bytes(b"not an index")
    with pytest.raises(OSError):
        save_from_file(gpu_backend(np.float32), str(tmp_path / "junk"))
    res.save.free()


def test_corrupt_headers_and_ids_are_rejected(tmp_path, oracle_mod):
    """ann_save_read does not trust the file: limits on every header field, par_maxes <= n, ids <= n,
    truncation (a bad d_short was a shift count, bad ids drove device reads)."""
    import struct
    from approximatenn_b200.api import gpu_backend, save_from_file, save_to_file
    g = load_golden("ragged_f32")
    orc = oracle_mod.restatement(g["dtype"])
    res = orc.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    lib = gpu_backend(np.float32)
    path = str(tmp_path / "index.annb")
    save_to_file(lib, res.save, path)
    good = open(path, "rb").read()
    tries, n, k = res.save.tries, res.save.n, res.save.k

    def rejected(blob):
        p = str(tmp_path / "bad.annb")
        open(p, "wb").write(blob)
        with pytest.raises(OSError):
            save_from_file(lib, p)

    def patched(offset, fmt, value):
        b = bytearray(good)
        struct.pack_into(fmt, b, offset, value)
        return bytes(b)

    save_from_file(lib, path).free()                              # the untouched file loads
    rejected(patched(12, "<I", 65))                                # tries
    rejected(patched(16, "<Q", 1 << 33))                           # n
    rejected(patched(24, "<Q", n))                                 # k >= n
    rejected(patched(32, "<Q", 64))                                # d_short as a shift count
    rejected(patched(48, "<Q", n + 1))                             # par_maxes[0] > n
    rejected(good[: len(good) // 2])                               # truncated
    graph_at = 48 + 8 * tries + 4 * res.save.d_long + 4 * tries * res.save.d_short * res.save.d_long
    rejected(patched(graph_at, "<I", n + 7))                       # a graph id beyond the sentinel
    res.save.free()
