"""Matrix Market text parsed on the device (csrc/cb_mmio.cu, cb_tile_from_mm_text): the lines of a share become the same triples
Python's float() / int() read out of them - the correctly rounded values strtod and the reference's sscanf give,
exact ties and 19-digit mantissas included
(include/CombBLAS/SpParMat.cpp:4010-4095, SpHelper.h:75-91) - with symmetric expansion, pattern files, duplicate merging, and
a clean refusal (CB_ERR_UNSUPPORTED, nothing exchanged) for numbers outside the parser's exact range."""
import numpy as np
import pytest

import cbb200_loader

cb = cbb200_loader.load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context(0)
    yield c
    c.close()


def tokens(rng, count):
    out = []
    for q in range(count):
        v = rng.standard_normal() * 10.0 ** int(rng.integers(-12, 13))
        out.append([f"{v:.17g}", f"{v:.6e}", f"{v:+.3f}", f"{int(v * 100) / 100:g}", f"{abs(v):.0f}.", f"0{abs(v):.2f}", f"{v:.10E}",
                    f"{v:.15g}", f"{rng.integers(-1000, 1000)}", f".{rng.integers(0, 10**9):09d}", f"{rng.integers(1, 2**53)}",
                    f"{rng.integers(1, 10**15)}e-22", f"{rng.integers(1, 10**15)}E+22", f"{(1 << 53) + 2 * int(rng.integers(0, 999)) + 1}",
                    f"{rng.integers(1, 10**18)}e{rng.integers(-280, 280)}", f"{v:.18e}"][q % 16])
    return out


def messy_text(m, n, cells, toks):
    lines = []
    for q, (c, t) in enumerate(zip(cells, toks)):
        sep, eol = ("\t", "\r\n") if q % 3 == 0 else ("  ", "\n") if q % 3 == 1 else (" ", " \n")
        lines.append(f"{c // n + 1}{sep}{c % n + 1}{sep}{t}{eol}" if t is not None else f"{c // n + 1}{sep}{c % n + 1}{eol}")
        if q % 50 == 7:
            lines.append("\n% a comment in the data section\n   \n")
    return "".join(lines).encode()


def tile_triples(t, dtype):
    rowptr, col, vals = t.to_csr(dtype)
    rows = np.repeat(np.arange(t.m), np.diff(rowptr))
    return rows, col, vals


@pytest.mark.parametrize("dtype,code", [(np.float64, "F64"), (np.float32, "F32"), (np.int32, "I32"), (np.int64, "I64")])
def test_values_are_what_strtod_reads(ctx, dtype, code):
    rng = np.random.default_rng(11)
    m, n = 613, 409
    cells = rng.choice(m * n, 20000, replace=False)
    toks = tokens(rng, len(cells))
    if dtype in (np.int32, np.int64):                       # keep the C cast defined: values inside the integer range
        toks = [t if abs(float(t)) < 2e9 else "7" for t in toks]
    t = ctx.tile_from_mm_text(m, n, messy_text(m, n, cells, toks), val_dtype=getattr(cb, code))
    rows, col, vals = tile_triples(t, dtype)
    order = np.argsort(cells)
    assert np.array_equal(rows * n + col, cells[order])
    with np.errstate(over="ignore"):                        # a double beyond the float range casts to inf, on both sides
        want = np.array([float(toks[q]) for q in order]).astype(dtype) if dtype in (np.float32, np.float64) else \
            np.array([int(float(toks[q])) for q in order], dtype)
    assert np.array_equal(vals.view(np.uint8), want.view(np.uint8))       # bit for bit, signed zeros included
    t.free()


def test_symmetric_pattern_zero_based_and_duplicates(ctx):
    rng = np.random.default_rng(3)
    n = 300
    I = rng.integers(0, n, 5000)
    J = rng.integers(0, n, 5000)
    lo = I >= J
    I, J = I[lo], J[lo]                                     # lower triangle with repeats
    # symmetric pattern file, zero-based
    text = "".join(f"{i} {j}\n" for i, j in zip(I, J)).encode()
    t = ctx.tile_from_mm_text(n, n, text, onebased=False, pattern=True, symmetric=True, val_dtype=cb.PATTERN)
    want = np.unique(np.concatenate([I * n + J, J * n + I]))
    rows, col, _ = tile_triples(t, None)
    assert np.array_equal(rows * n + col, want)
    t.free()
    # values: duplicates merged by sum / max / min / first
    V = rng.integers(-50, 50, len(I)).astype(np.float64)
    text = "".join(f"{i + 1} {j + 1} {v:g}\n" for i, j, v in zip(I, J, V)).encode()
    key = I * n + J
    uk, inv = np.unique(key, return_inverse=True)
    for op, red in [(1, np.add), (2, np.maximum), (3, np.minimum)]:
        acc = np.full(len(uk), {1: 0.0, 2: -np.inf, 3: np.inf}[op])
        red.at(acc, inv, V)
        t = ctx.tile_from_mm_text(n, n, text, val_dtype=cb.F64, dup_op=op)
        rows, col, vals = tile_triples(t, np.float64)
        assert np.array_equal(rows * n + col, uk) and np.array_equal(vals, acc)
        t.free()
    first = np.full(len(uk), np.nan)
    for k, v in zip(inv[::-1], V[::-1]):
        first[k] = v
    t = ctx.tile_from_mm_text(n, n, text, val_dtype=cb.F64, dup_op=0)
    assert np.array_equal(tile_triples(t, np.float64)[2], first)
    t.free()


def test_empty_share_and_lines_without_entries(ctx):
    t = ctx.tile_from_mm_text(5, 5, b"", val_dtype=cb.F64)
    assert t.nnz == 0
    t.free()
    t = ctx.tile_from_mm_text(5, 5, b"\n\n% nothing\n   \r\n3 4 2.5", val_dtype=cb.F64)          # last line without a newline
    rows, col, vals = tile_triples(t, np.float64)
    assert (rows.tolist(), col.tolist(), vals.tolist()) == ([2], [3], [2.5])
    t.free()


@pytest.mark.parametrize("bad", ["0.1234567890123456789012345", "1e400", "inf", "nan", "0x1p3", "12345678901234567890", "4.9e-324", "1.5f", ""])
def test_numbers_outside_the_exact_range_are_refused(ctx, bad):
    with pytest.raises(cb.CBError) as e:
        ctx.tile_from_mm_text(5, 5, f"1 1 1.0\n2 2 {bad}\n".encode(), val_dtype=cb.F64)
    assert e.value.status == 4                              # CB_ERR_UNSUPPORTED: the host layer then parses the file on the CPU


def test_entry_outside_the_matrix_is_an_error(ctx):
    with pytest.raises(cb.CBError) as e:
        ctx.tile_from_mm_text(5, 5, b"6 1 1.0\n", val_dtype=cb.F64)
    assert e.value.status == 3007


def test_large_share(ctx):
    # two million lines: what a rank sees of a file of a few hundred MB on 8 ranks
    rng = np.random.default_rng(2)
    n = 1 << 20
    nz = 2_000_000
    I = rng.integers(0, n, nz)
    J = rng.integers(0, n, nz)
    V = rng.integers(1, 10**6, nz)
    text = "\n".join(f"{i + 1} {j + 1} {v}.5e-3" for i, j, v in zip(I.tolist(), J.tolist(), V.tolist())).encode()
    t = ctx.tile_from_mm_text(n, n, text, val_dtype=cb.F64, dup_op=2)
    key = I * n + J
    order = np.lexsort((V, key))
    ks, vs = key[order], V[order]
    last = np.r_[ks[1:] != ks[:-1], True]
    rows, col, vals = tile_triples(t, np.float64)
    assert np.array_equal(rows * n + col, ks[last])
    assert np.array_equal(vals, np.array([float(f"{v}.5e-3") for v in vs[last].tolist()]))
    t.free()
