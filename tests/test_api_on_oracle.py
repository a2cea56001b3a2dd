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


# ---- the reference's own engine-dependent tests (tests/test_main_pytest.py), same names; --taxa Metazoa stands in for Fungi
# (F.hmm is missing from the mount), so the record counts are this taxon's, not the reference's 235 / 226 ----
import os  # noqa: E402

from conftest import TD  # noqa: E402


def _cli(tmp_path, *argv):
    from itsxpress_b200 import fastq as fq
    from itsxpress_b200 import main as cli
    out = str(tmp_path / "testout.fastq")
    cli.main(args=cli.myparser().parse_args(list(argv) + ["--outfile", out, "--region", "ITS2", "--taxa", "Metazoa", "--threads", "1",
                                                          "--log", str(tmp_path / "log.txt"), "--tempdir", str(tmp_path)]))
    return fq.read_fastq(out).n


def test_seq_sample_not_paired(on_oracle, tmp_path):                       # reference :165-171
    from itsxpress_b200.main import create_runtime_hmm
    sobj = SeqSample.SeqSampleNotPaired(fastq=os.path.join(TD, "4774-1-MSITS3_merged.fastq"), tempdir=str(tmp_path))
    sobj.deduplicate(threads=1)
    sobj._search(hmmfile=create_runtime_hmm("Metazoa", "ITS2", str(tmp_path)), threads=1)
    assert os.path.getsize(sobj.dom_file) > 10000 and os.path.getsize(sobj.uc_file) > 1000


def test_seq_sample_paired_not_interleaved(on_oracle, tmp_path):           # reference :183-193
    from itsxpress_b200.main import create_runtime_hmm
    sobj = SeqSample.SeqSamplePairedNotInterleaved(fastq=os.path.join(TD, "4774-1-MSITS3_R1.fastq"), tempdir=str(tmp_path),
                                                   fastq2=os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    sobj._merge_reads(stagger=True, threads=1)
    sobj.deduplicate(threads=1)
    sobj._search(hmmfile=create_runtime_hmm("Metazoa", "ITS2", str(tmp_path)), threads=1)
    assert os.path.getsize(sobj.seq_file) > 100000 and os.path.getsize(sobj.dom_file) > 10000


def test_main_paired(on_oracle, tmp_path):                                  # reference :226-253 (235 with Fungi)
    n = _cli(tmp_path, "--fastq", os.path.join(TD, "4774-1-MSITS3_R1.fastq"), "--fastq2", os.path.join(TD, "4774-1-MSITS3_R2.fastq"))
    assert 150 < n <= 250


def test_main_paired_high_qual(on_oracle, tmp_path):                        # reference :256-285: quality scores above 41
    """The reference asserts nothing here but that the run ends (its count is commented out): most base qualities of this
    pair file were overwritten with Q49, so nearly every overlap scores below the merger's minimum (4 of 250 pairs merge
    in the restatement) -- what the test guards is --fastq_qmax 93: qualities above 41 are not an error."""
    n = _cli(tmp_path, "--fastq", os.path.join(TD, "high_qual_scores_R1.fastq.gz"), "--fastq2",
             os.path.join(TD, "high_qual_scores_R2.fastq.gz"))
    assert 0 <= n <= 250 and "merge_pairs" in on_oracle.calls


def test_main_merged(on_oracle, tmp_path):                                  # reference :288-314 (226 with Fungi)
    n = _cli(tmp_path, "--fastq", os.path.join(TD, "4774-1-MSITS3_merged.fastq"), "--single_end")
    assert 150 < n <= 227


def test_main_paired_no_cluster(on_oracle, tmp_path):                       # reference :317-347: --cluster_id 1
    n = _cli(tmp_path, "--fastq", os.path.join(TD, "4774-1-MSITS3_R1.fastq"), "--fastq2", os.path.join(TD, "4774-1-MSITS3_R2.fastq"),
             "--cluster_id", "1")
    assert 150 < n <= 250
