"""The reference's own known-answer tests, replayed through the public API of otters_b200 (CUDA path)."""
import pytest

from helpers import check_expect, check_meta_expect, load_kats, product_meta_kat, product_vec_kat

KATS = load_kats()
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kat", KATS["vec"], ids=[k["name"] for k in KATS["vec"]])
def test_vecstore_kats(kat, ctx):
    check_expect(kat, product_vec_kat(kat))


@pytest.mark.parametrize("kat", KATS["meta"], ids=[k["name"] for k in KATS["meta"]])
def test_metastore_kats(kat, ctx):
    result, stats = product_meta_kat(kat)
    check_meta_expect(kat, result, stats)
