"""Pins the CPU oracle on the reference's own end-to-end fixtures."""
import pytest

from golden_check import check_scenario

# scenarios whose checked outputs this harness regenerates (split / annotate / combine / ga4gh / roc-only modes): all 29 bundled fixtures
SCENARIOS = ["small", "small_squash", "small_alt", "tricky", "tricky2", "tricky3", "tricky4", "obeyphasing",
             "small_region", "small_samples", "too_complex", "too_complex_end", "updel", "updel2", "updel3", "updel4",
             "updel5", "trickysquash", "tetraploid", "annotate", "roc_only", "split", "split_no_roc",
             # --output-mode combine: the two-sample output.vcf line by line (oracle/ref_host.py format_combine)
             "combine", "combine2", "triploid", "avoid_overlap",
             # --output-mode ga4gh: TRUTH / QUERY columns with BD / BK / BI behind the loose-match filter (format_ga4gh)
             "ga4gh", "ga4gh2"]


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_matches_reference_golden(oracle, fixtures, name):
    checked = check_scenario(oracle, fixtures[name])
    assert checked, "nothing was compared for " + name
