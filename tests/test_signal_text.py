"""CPU: `.signal` text ingestion (utils/labelop.py:199-219 reads the same int16 `Signal` samples from fast5): the integer
parser in libnanodec against numpy's float parser, and the fallbacks / refusals of read_raw_signal."""
import numpy as np
import pytest

from nanodecoder_b200.utils.labelop import read_raw_signal


def _write(tmp_path, text, name="r.signal"):
    p = tmp_path / name
    p.write_bytes(text.encode())
    return str(p)


def test_integer_samples_match_numpy(tmp_path):
    rng = np.random.default_rng(3)
    x = rng.integers(-32768, 32768, size=50000)
    x[:4] = [-32768, 32767, 0, -1]
    for sep in (" ", "\n", "\r\n", " \t "):
        got = read_raw_signal(_write(tmp_path, sep.join(map(str, x)) + sep), "signal")
        assert got.dtype == np.int16
        np.testing.assert_array_equal(got, x.astype(np.int16))
    got = read_raw_signal(_write(tmp_path, "+5 -0 007"), "signal")
    np.testing.assert_array_equal(got, [5, 0, 7])
    assert read_raw_signal(_write(tmp_path, ""), "signal").size == 0
    assert read_raw_signal(_write(tmp_path, " \n "), "signal").size == 0


def test_float_formatted_integers_take_the_general_parser(tmp_path):
    got = read_raw_signal(_write(tmp_path, "512.0 4.81e2 -3"), "signal")
    np.testing.assert_array_equal(got, [512, 481, -3])


def test_float_valued_samples_come_back_as_float64(tmp_path):
    """The reference takes any float token (utils/labelop.py:216-217); non-integer or out-of-int16 values go to the fp64
    front-end kernels as float64, parsed by Python's float() like the reference does."""
    got = read_raw_signal(_write(tmp_path, "1 2.5 3"), "signal")
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, [1.0, 2.5, 3.0])
    got = read_raw_signal(_write(tmp_path, "1 40000 -3e-7"), "signal")
    assert got.dtype == np.float64 and got[1] == 40000.0 and got[2] == float("-3e-7")
    with pytest.raises(ValueError):
        read_raw_signal(_write(tmp_path, "12a 3"), "signal")              # float("12a") fails in the reference too
