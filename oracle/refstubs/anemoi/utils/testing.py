"""anemoi.utils.testing for the reference's test-suite when it runs offline: test data cannot
be fetched, so `skip_if_offline` always skips and `get_test_data` skips the test that asks."""
import pytest

skip_if_offline = pytest.mark.skip(reason="offline: no test data")
skip_slow_tests = pytest.mark.skip(reason="slow")
skip_missing_packages = lambda *a, **k: pytest.mark.skip(reason="missing packages")  # noqa: E731


class GetTestData:
    def __call__(self, path, **kwargs):
        pytest.skip(f"offline: cannot fetch {path}")


@pytest.fixture
def get_test_data():
    return GetTestData()
