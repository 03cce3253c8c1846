"""The CPU oracle against the golden fixtures (outputs of the unmodified reference, tests/golden/).
Runs anywhere — no GPU, no /root/reference."""
import pytest

from refutil import assert_rows_equal, golden_names, load_golden, oracle_regions, oracle_strand_shift


@pytest.mark.parametrize("name", golden_names("regions"))
def test_oracle_regions_golden(name):
    data, st, want = load_golden(name)
    got = oracle_regions(data, st)
    assert got["status"] == 0
    # the reference's putRegion evaluates strandCorr() for every row; the oracle does too
    assert_rows_equal(got["rows"], want["rows"], name + " written")
    assert_rows_equal(got["dropped"], want["dropped"], name + " dropped")
    assert got["n_passed"] == want["n_passed"] and got["n_rejected"] == want["n_rejected"]
    assert got["tags_in_regions"] == want["tags_in_regions"]


@pytest.mark.parametrize("name", golden_names("shift"))
def test_oracle_shift_golden(name):
    data, p, want = load_golden(name)
    got = oracle_strand_shift(data, **p)
    assert got["status"] == 0
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        assert got[k] == want[k], k
    assert got["background"] == want["background"]
    assert got["tested"] == want["tested"]


def test_fixture_set_is_complete():
    names = golden_names("regions")
    for needed in ("regions_directional", "regions_nondirectional", "regions_controls_q4", "regions_coeffs_q3", "regions_left_edge_q2",
                   "regions_one_contig_q7", "regions_long_q8"):
        assert needed in names


@pytest.mark.parametrize("name", golden_names("tir"))
def test_oracle_tags_in_regions_golden(name):
    from refutil import oracle_tags_in_regions
    data, p, want = load_golden(name)
    status, got = oracle_tags_in_regions(data, p["keys"], directional=p["directional"], extension=p["extension"])
    assert status == 0
    assert got.tolist() == want.tolist()
