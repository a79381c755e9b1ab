"""MLUtilsSuite-style LIBSVM cases (mllib test MLUtilsSuite.scala:82-150)."""
import os

import numpy as np
import pytest
import torch

from asyncframework_b200.data import load_libsvm_file, parse_libsvm_record, save_as_libsvm_file
from asyncframework_b200.data.libsvm import load_libsvm_csr, MLUtils


def _write(tmp_path, text, name="part-00000"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_load_basic_with_blank_comment_and_label_only(tmp_path, scheduler):
    path = _write(tmp_path, "1 1:1.0 3:2.0 5:3.0\n\n# a comment\n0\n0 2:4.0 4:5.0 6:6.0\n   \n")
    labels, indptr, indices, values, mx = load_libsvm_csr(path)
    assert labels.tolist() == [1.0, 0.0, 0.0] and indptr.tolist() == [0, 3, 3, 6]
    assert indices.tolist() == [0, 2, 4, 1, 3, 5] and values.tolist() == [1, 2, 3, 4, 5, 6] and mx == 5
    ds = load_libsvm_file(scheduler, path, num_partitions=2)
    assert ds.num_features == 6 and ds.count() == 3
    recs = ds.collect()
    assert recs[1].features.num_nonzeros == 0                    # label-only line -> all-zero sparse row
    assert recs[0].features.to_dense().values.tolist() == [1, 0, 2, 0, 3, 0]
    ds8 = load_libsvm_file(scheduler, path, num_features=8)
    assert ds8.num_features == 8


def test_zero_based_and_descending_indices_rejected(tmp_path):
    with pytest.raises(ValueError, match="one-based"):
        load_libsvm_csr(_write(tmp_path, "1 1:1 3:2\n0 0:1 2:3\n", "a"))
    with pytest.raises(ValueError, match="ascending"):
        load_libsvm_csr(_write(tmp_path, "1 3:1 2:2\n", "b"))
    with pytest.raises(ValueError):
        parse_libsvm_record("1 3:1 3:2")
    with pytest.raises(ValueError):
        load_libsvm_csr(_write(tmp_path, "1 2:x\n", "c"))


def test_multiple_spaces_and_python_parser_agree(tmp_path):
    line = "-1.5  2:0.5   7:1e-3 9:-2"
    lab, idx, val = parse_libsvm_record(line)
    labels, indptr, indices, values, mx = load_libsvm_csr(_write(tmp_path, line + "\n"))
    assert lab == labels[0] == -1.5 and idx.tolist() == indices.tolist() == [1, 6, 8]
    assert np.allclose(val, values)


def test_save_load_roundtrip_directory(tmp_path, scheduler):
    rng = np.random.RandomState(0)
    X = rng.randn(37, 9) * (rng.rand(37, 9) < 0.4)
    y = rng.randn(37)
    from asyncframework_b200.data import ShardedDataset
    ds = ShardedDataset.from_arrays(scheduler, X, y, 3)
    out = str(tmp_path / "out")
    save_as_libsvm_file(ds, out)
    assert sorted(os.listdir(out)) == ["_SUCCESS", "part-00000", "part-00001", "part-00002"]
    back = load_libsvm_file(scheduler, out, num_features=9, num_partitions=3)    # directory input
    rows = {round(float(r.label), 12): r.features.to_dense().values for r in back.collect()}
    for i in range(37):
        assert torch.allclose(rows[round(float(y[i]), 12)], torch.tensor(X[i]), atol=1e-12)
    with pytest.raises(FileExistsError):
        save_as_libsvm_file(ds, out)


def test_large_file_multithreaded_matches_single_thread(tmp_path):
    rng = np.random.RandomState(1)
    lines = []
    for i in range(20000):
        k = rng.randint(0, 12)
        idx = np.sort(rng.choice(500, size=k, replace=False)) + 1
        lines.append(" ".join([str(i % 3)] + [f"{j}:{rng.randn():.6f}" for j in idx]))
    path = _write(tmp_path, "\n".join(lines) + "\n")
    from asyncframework_b200.data.libsvm import parse_libsvm_file_csr
    a = parse_libsvm_file_csr(path, 1)
    b = parse_libsvm_file_csr(path, 8)
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    assert a[4] == b[4]


def test_mlutils_helpers():
    from asyncframework_b200.linalg import Vectors
    v = MLUtils.append_bias(Vectors.dense(1.0, 2.0))
    assert v.values.tolist() == [1.0, 2.0, 1.0]
    s = MLUtils.append_bias(Vectors.sparse(3, [1], [4.0]))
    assert s.size == 4 and s.indices.tolist() == [1, 3]
    import math
    assert abs(MLUtils.log1p_exp(800.0) - 800.0) < 1e-9 and abs(MLUtils.log1p_exp(-800.0)) < 1e-9
    assert abs(MLUtils.log1p_exp(0.3) - math.log(1 + math.exp(0.3))) < 1e-12


def test_native_decimal_fast_path_is_correctly_rounded(tmp_path):
    """The parser converts plain decimals without strtod; every value must equal Python's float(token) bit for bit,
    including tokens that must take the slow path (exponents, > 15 digits, inf/nan)."""
    import random
    from asyncframework_b200.data.libsvm import parse_libsvm_file_csr
    rng = random.Random(7)
    toks = ["0", "-0", "+1", "1.", ".5", "-.25", "0.1", "0.30000000000000004", "123456789012345", "9007199254740993", "1e-3", "-2.5E+7",
            "0.000000000000000000001", "123456.789012345678901234", "1e22", "1e23", "4.9e-324", "1.7976931348623157e308", "inf", "-inf",
            "00012.5000", "3.14159265358979323846", "255", "0.00392156862745098", "179769313486231570000"]
    for _ in range(3000):
        nd, nf = rng.randint(0, 17), rng.randint(0, 20)
        t = "".join(rng.choice("0123456789") for _ in range(nd))
        if nf:
            t += "." + "".join(rng.choice("0123456789") for _ in range(nf))
        if not any(c.isdigit() for c in t):
            t = "7" + t
        toks.append(rng.choice(["", "-", "+"]) + t)
    path = tmp_path / "vals.libsvm"
    with open(path, "w") as f:
        for i in range(0, len(toks), 10):
            row = toks[i:i + 10]
            f.write(f"{row[0]} " + " ".join(f"{j + 1}:{v}" for j, v in enumerate(row)) + "\n")
    labels, indptr, indices, values, mx = parse_libsvm_file_csr(str(path), num_threads=3)
    exp = np.array([float(t) for t in toks], dtype=np.float64)
    assert len(values) == len(toks)
    assert np.array_equal(values.view(np.uint64), exp.view(np.uint64))
    assert np.array_equal(labels.view(np.uint64), exp[::10].view(np.uint64))
    assert np.array_equal(np.diff(indptr), np.minimum(10, len(toks) - np.arange(0, len(toks), 10)))


@pytest.mark.parametrize("bad", ["1 1:2 x:3", "1 2:1 2:3", "1 0:5", "1 3:", "1 :4", "1 3:4:5", "abc 1:2", "1 1:2\t2:3", "1 5:1e", "1 1:--2"])
def test_native_parser_rejects_malformed_tokens(tmp_path, bad):
    from asyncframework_b200.data.libsvm import parse_libsvm_file_csr
    path = tmp_path / "bad.libsvm"
    path.write_text("1 1:1\n" + bad + "\n")
    with pytest.raises(ValueError):
        parse_libsvm_file_csr(str(path))


def test_native_parser_rejects_indices_beyond_int32_and_reports_absolute_line_numbers(tmp_path):
    """A feature index above 2^31 must fail like MLlib's `toInt` (not wrap to a small / negative column), and parse errors
    carry the line number in the FILE even when the file is parsed in parallel chunks."""
    from asyncframework_b200.data.libsvm import parse_libsvm_file_csr as load_libsvm_file_native
    good = "1 1:0.5 3:1.5\n"
    p = tmp_path / "big.libsvm"
    p.write_text(good * 4000 + "0 2:1.0 3000000000:2.0\n" + good * 10)
    with pytest.raises(ValueError) as ei:
        load_libsvm_file_native(str(p), num_threads=4)
    assert "32-bit" in str(ei.value) and "line 4001" in str(ei.value)
    q = tmp_path / "order.libsvm"
    q.write_text(good * 7000 + "1 5:1.0 2:1.0\n")
    with pytest.raises(ValueError) as ei:
        load_libsvm_file_native(str(q), num_threads=4)
    assert "ascending" in str(ei.value) and "line 7001" in str(ei.value)
