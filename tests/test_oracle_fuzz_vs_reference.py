"""The randomised workloads of test_zz_fuzz_gpu.py, oracle against the compiled, unmodified reference (ref_replay: raw doubles bit for
bit): the fuzz cases the GPU is later held to are themselves pinned to the reference."""
import os

import pytest

import refutil
from refutil import oracle_regions, region_settings, run_ref_replay_regions
from test_oracle_vs_reference import assert_same_rows
from test_zz_fuzz_gpu import _case

pytestmark = pytest.mark.skipif(not refutil.have_ref(), reason="compiled reference (oracle/_ref) not present")


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("UPB_FUZZ_SEEDS", "96")))))
def test_random_workload_oracle_equals_reference(tmp_path, seed):
    d, kw = _case(seed)
    if sum(d.tags(s) for s in range(d.n_samples) if s not in kw.get("control", [])) == 0:
        pytest.skip("no non-control tags left (the reference divides by their number)")
    st = region_settings(d, **kw)
    ours = oracle_regions(d, st)
    if ours["status"] != 0:
        with pytest.raises(RuntimeError):                               # the reference trips its own assert (Q7)
            run_ref_replay_regions(d, st, str(tmp_path))
        return
    ref = run_ref_replay_regions(d, st, str(tmp_path))
    assert_same_rows(ours["rows"], ref["rows"], "seed %d written" % seed)
    assert_same_rows(ours["dropped"], ref["dropped"], "seed %d dropped" % seed)
    assert ours["n_passed"] == ref["n_passed"] and ours["n_rejected"] == ref["n_rejected"]
    assert ours["tags_in_regions"] == ref["tags_in_regions"]


@pytest.mark.parametrize("seed", list(range(32)))
def test_random_strand_shift_oracle_equals_reference(tmp_path, seed):
    from refutil import oracle_strand_shift, run_ref_replay_shift
    from test_zz_fuzz_gpu import _shift_case
    d, kw = _shift_case(seed)
    ours = oracle_strand_shift(d, **kw)
    ref = run_ref_replay_shift(d, str(tmp_path), **kw)
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        assert ours[k] == ref[k], (seed, k)
    for g, w in zip(ours["tested"], ref["tested"]):
        assert (g["contig"], g["left"], g["len"], g["best_shift"]) == (w["contig"], w["left"], w["len"], w["best_shift"])
        assert g["best_corr"] == w["best_corr"] or (g["best_corr"] != g["best_corr"] and w["best_corr"] != w["best_corr"])
