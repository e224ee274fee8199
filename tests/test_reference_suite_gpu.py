"""The reference's own test-suite (tests/test_wavelets.py, tests/test_utils.py of watroo 0.0.4), re-stated against
wavelets_b200 with the same names and assertions: a user switching the import line keeps a green suite."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def data_2d():
    # tests/__init__.py of the reference: a constant 128 x 128 image
    return np.ones((128, 128))


class TestTransform:

    def test_regular(self, data_2d):
        from wavelets_b200 import AtrousTransform
        transform = AtrousTransform()
        regular = transform(data_2d, 4)
        expected = np.zeros(regular.data.shape)  # first planes should be zeros
        expected[-1] = 1  # last planes should be ones
        assert np.isclose(regular, expected).all()

    def test_regular_vs_recursive(self, data_2d):
        from wavelets_b200 import AtrousTransform
        transform = AtrousTransform()
        regular = transform(data_2d, 4)
        recursive = transform(data_2d, 4, recursive=True)
        assert np.isclose(regular, recursive).all()


class TestWOW:

    def test_wow(self, data_2d):
        from wavelets_b200.utils import wow
        wowed, _ = wow(data_2d)
        assert wowed.shape == data_2d.shape
        wowed, _ = wow(data_2d, bilateral=True)
        assert wowed.shape == data_2d.shape
