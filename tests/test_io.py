"""CPU suite, SURVEY 8f row f4: the callers' snapshot files -- "x y w" text, j outer / i inner, Julia's own Float64
printing (vm.jl:81-85,132-136,142-146; reader plotting.jl:14-28) -- written and read by the library's native code
(csrc/vmk_io.hpp).  Host-side I/O: no device is involved, so the product library itself is exercised here."""
import os
from decimal import Decimal

import numpy as np
import pytest

from helpers import grid, vm_field

# Julia's print(::Float64) for a few values (Julia >= 1.6, Ryu `writeshortest`): positional for 1e-4 <= |v| < 1e6,
# otherwise d.ddde[-]X with neither '+' nor zero padding in the exponent
JULIA_PRINT = {
    1.0: "1.0", 0.1: "0.1", 1e-5: "1.0e-5", 1e-4: "0.0001", 0.001: "0.001", 0.00012345: "0.00012345",
    123456.7: "123456.7", 1e5: "100000.0", 999999.0: "999999.0", 1e6: "1.0e6", 1234567.0: "1.234567e6",
    12345678.9: "1.23456789e7", 6.283185307179586: "6.283185307179586", 0.04908738521234052: "0.04908738521234052",
    0.1 + 0.2: "0.30000000000000004", 5e-324: "5.0e-324", 1.7976931348623157e308: "1.7976931348623157e308",
    1e22: "1.0e22", 1e15: "1.0e15", 1e16: "1.0e16", 2.5e-7: "2.5e-7", -3.25: "-3.25", -1e-10: "-1.0e-10", 3.0: "3.0",
    1.5e300: "1.5e300", 0.0: "0.0", float("inf"): "Inf", float("-inf"): "-Inf",
}


@pytest.fixture(scope="module")
def lib():
    import cfd_julia_b200
    return cfd_julia_b200.default_library()  # libvmk.so: loads without a GPU, the I/O entry points are host code


def julia_print_restated(v: float) -> str:
    """Independent restatement of base/ryu/shortest.jl's layout on top of Python's shortest round-trip digits."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Inf" if v > 0 else "-Inf"
    sign = "-" if np.signbit(v) else ""
    v = abs(v)
    if v == 0:
        return sign + "0.0"
    _, digits, exp = Decimal(repr(v)).as_tuple()
    d = "".join(map(str, digits)).rstrip("0") or "0"
    exp += len(digits) - len(d)
    pt = len(d) + exp
    if -4 < pt <= 6:
        if pt <= 0:
            return sign + "0." + "0" * (-pt) + d
        if pt < len(d):
            return sign + d[:pt] + "." + d[pt:]
        return sign + d + "0" * (pt - len(d)) + ".0"
    return sign + d[0] + "." + (d[1:] or "0") + "e" + str(pt - 1)


def test_julia_float_printing_known_values(lib):
    from cfd_julia_b200.common import julia_float_str
    for v, s in JULIA_PRINT.items():
        assert julia_float_str(v, lib) == s, (v, s)
        assert julia_print_restated(v) == s
    assert julia_float_str(float("nan"), lib) == "NaN"
    assert julia_float_str(-0.0, lib) == "-0.0" == julia_print_restated(-0.0)


def test_julia_float_printing_random(lib):
    """shortest round trip: identical digits to an independent shortest-digits source, and float(s) == v bit for bit"""
    from cfd_julia_b200.common import julia_float_str
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2**63, 4000, dtype=np.uint64) | (rng.integers(0, 2, 4000, dtype=np.uint64) << np.uint64(63))
    vals = list(bits.view(np.float64)) + list(rng.uniform(-7, 7, 1000)) + list(10.**rng.uniform(-8, 8, 1000))
    for v in vals:
        v = float(v)
        if v != v or abs(v) == float("inf"):
            continue
        s = julia_float_str(v, lib)
        assert s == julia_print_restated(v), (v, s)
        assert float(s) == v and (np.signbit(float(s)) == np.signbit(v))


@pytest.mark.parametrize("n", [16, 64])
def test_write_read_field(lib, tmp_path, n):
    from cfd_julia_b200.common import read_field, write_field
    dx, dy, x, y = grid(n)
    ut = np.asfortranarray(vm_field(n)[1:n + 2, 1:n + 2])
    ut[3, 5] = -1.25e-7  # exponent form, negative
    ut[0, 1] = 0.
    path = str(tmp_path / "vm0.txt")
    write_field(path, x, y, ut, lib=lib)
    text = open(path).read()
    lines = text.split("\n")
    assert lines[-1] == "" and len(lines) == (n + 1)**2 + 1
    want = [f"{julia_print_restated(float(x[i]))} {julia_print_restated(float(y[j]))} {julia_print_restated(float(ut[i, j]))}"
            for j in range(n + 1) for i in range(n + 1)]  # vm.jl:82-84: j outer, i inner
    assert lines[:-1] == want
    assert lines[0].startswith("0.0 0.0 ")
    xx, yy, w = read_field(path, n, n, lib=lib)  # plotting.jl:14-28
    assert np.array_equal(xx, x) and np.array_equal(yy, y) and np.array_equal(w, ut)  # shortest digits round-trip exactly


def test_io_errors(lib, tmp_path):
    from cfd_julia_b200.common import VmkError, read_field, write_field
    x = np.arange(5.)
    with pytest.raises(VmkError) as e:
        write_field(str(tmp_path / "no_such_dir" / "f.txt"), x, x, np.zeros((5, 5), order="F"), lib=lib)
    assert e.value.code == 5
    with pytest.raises(IndexError):
        write_field(str(tmp_path / "f.txt"), x[:3], x, np.zeros((5, 5), order="F"), lib=lib)
    with pytest.raises(VmkError):
        read_field(str(tmp_path / "missing.txt"), 4, 4, lib=lib)
    bad = tmp_path / "bad.txt"
    bad.write_text("0.0 0.0 1.0\n0.1 oops 2.0\n")
    with pytest.raises(VmkError):
        read_field(str(bad), 1, 0, lib=lib)
    short = tmp_path / "short.txt"
    short.write_text("0.0 0.0 1.0\n\n1.0 0.0 2.0\n")  # readdlm skips the blank line; 2 rows != 2*2
    with pytest.raises(ValueError):
        read_field(str(short), 1, 1, lib=lib)
    xx, yy, w = read_field(str(short), 1, 0, lib=lib)
    assert list(xx) == [0., 1.] and list(yy) == [0.] and w.shape == (2, 1) and list(w[:, 0]) == [1., 2.]


def test_numerical_snapshot_files_are_numbered(lib):
    """vm.jl:21,78-86 never increments its record index (every snapshot overwrites vm1.txt while plotting.jl:14,26 reads
    vm5.txt / vm10.txt); the mirror numbers the files vm1 .. vm{ns} -- checked on the file-name logic only (no device)."""
    import inspect

    from cfd_julia_b200 import common
    src = inspect.getsource(common.Common._numerical)
    assert "rec[0] += 1" in src and 'vm{rec[0]}.txt' in src
