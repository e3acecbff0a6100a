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


def test_non_integer_or_out_of_range_samples_are_refused(tmp_path):
    with pytest.raises(ValueError):
        read_raw_signal(_write(tmp_path, "1 2.5 3"), "signal")        # normalised floats are not raw DAC samples
    with pytest.raises(ValueError):
        read_raw_signal(_write(tmp_path, "1 40000 3"), "signal")
    with pytest.raises(ValueError):
        read_raw_signal(_write(tmp_path, "1 99999999999999999999 3"), "signal")
    with pytest.raises(ValueError):
        read_raw_signal(_write(tmp_path, "12a 3"), "signal")
