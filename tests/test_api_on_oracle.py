"""The reference-facing tests of the GPU suite (tests/test_gpu_api.py, and the host-level ones of tests/test_gpu_merge.py)
run a second time on the CPU with the oracle standing in for the device (tests/oracle_context.py): the SAME test bodies --
Dedup / ItsPosition / SeqSample / the command line / the QIIME 2 actions against the reference's golden files (QIIME 2
single-end golden bytes, t2_r1.fq / t2_r2.fq, vsearch's uc.txt / rep.fa) -- so that the host code above the C ABI is held
to the reference's fixtures by the `-m "not gpu"` suite too, and a regression in it shows up before the round-end GPU run.
What stays GPU-only: everything that calls the library directly (kernels against the oracle)."""
import pytest

import test_gpu_api as G
import test_gpu_merge as M
from itsxpress_b200 import SeqSample
from oracle_context import OracleContext


@pytest.fixture()
def on_oracle(oracle, monkeypatch):
    ctx = OracleContext(oracle)
    monkeypatch.setattr(SeqSample, "get_context", lambda: ctx)
    SeqSample.reset_sessions()
    yield ctx
    SeqSample.reset_sessions()


def test_deduplicate_writes_vsearch_files(on_oracle, tmp_path):
    G.test_deduplicate_writes_vsearch_files(tmp_path)


@pytest.mark.parametrize("mode", ["plain", "gz", "zst"])
def test_create_trimmed_seqs_golden(on_oracle, tmp_path, mode):
    G.test_create_trimmed_seqs_golden(tmp_path, mode)


def test_create_paired_trimmed_seqs_golden(on_oracle, tmp_path):
    G.test_create_paired_trimmed_seqs_golden(tmp_path)


def test_trim_ccs_bulk(on_oracle, tmp_path):
    G.test_trim_ccs_bulk(tmp_path)


def test_pipeline_objects_device_vs_files(on_oracle, tmp_path, oracle):
    G.test_pipeline_objects_device_vs_files(tmp_path, oracle)


def test_cli_end_to_end(on_oracle, tmp_path, oracle, caplog):
    G.test_cli_end_to_end(tmp_path, oracle, caplog)


def test_cli_failure_exit_code(on_oracle, tmp_path):
    G.test_cli_failure_exit_code(tmp_path)


def test_search_without_profiles_fails_like_hmmsearch(on_oracle, tmp_path):
    G.test_search_without_profiles_fails_like_hmmsearch(tmp_path)


def test_q2_trim_single(on_oracle, tmp_path):
    G.test_q2_trim_single(tmp_path)


def test_pipeline_without_materialised_temp_files(on_oracle, tmp_path):
    G.test_pipeline_without_materialised_temp_files(tmp_path)


def test_merge_reads_api_writes_seq_fq(on_oracle, tmp_path, oracle):
    M.test_merge_reads_api_writes_seq_fq(oracle, tmp_path)


def test_merge_reads_errors(on_oracle, tmp_path):
    M.test_merge_reads_errors(tmp_path)


def test_q2_trim_pair_actions(on_oracle, tmp_path, monkeypatch):
    from itsxpress_b200 import q2_itsxpress as q2
    monkeypatch.setattr(q2, "BATCH_READS", 0)
    M.test_q2_trim_pair_actions(tmp_path)


def test_q2_main_sharded_single_rank_equals_action(on_oracle, tmp_path):
    M.test_q2_main_sharded_single_rank_equals_action(tmp_path)
