"""Constants shared across the package (same names and values as the reference's itsxpress/definitions.py:7,32-82,
which the CLI and the QIIME 2 plugin import).  Data, not logic: the taxon -> profile-file table of ITSx_db."""
import os

ROOT_DIR = os.path.dirname(os.path.abspath(__file__))

# (taxon name as the CLI accepts it, ITSx_db file).  Order matters: `--taxa All` searches the files in this
# order (main.py:192-195).  The leading blank of " Rhizaria" is the reference's own key (definitions.py:47,70)
# and is kept so that the command line accepts exactly the same strings.
_TAXA = (
    ("Alveolata", "A.hmm"), ("Bryophyta", "B.hmm"), ("Bacillariophyta", "C.hmm"), ("Amoebozoa", "D.hmm"),
    ("Euglenozoa", "E.hmm"), ("Fungi", "F.hmm"), ("Chlorophyta", "G.hmm"), ("Rhodophyta", "H.hmm"),
    ("Phaeophyceae", "I.hmm"), ("Marchantiophyta", "L.hmm"), ("Metazoa", "M.hmm"), ("Oomycota", "O.hmm"),
    ("Haptophyceae", "P.hmm"), ("Raphidophyceae", "Q.hmm"), (" Rhizaria", "R.hmm"), ("Synurophyceae", "S.hmm"),
    ("Tracheophyta", "T.hmm"), ("Eustigmatophyceae", "U.hmm"), ("Parabasalia", "Y.hmm"), ("All", "all.hmm"),
)
taxa_choices = [t for t, _ in _TAXA]
taxa_dict = dict(_TAXA)

maxmismatches = 40          # vsearch --fastq_maxdiffs (merge step, outside the GPU path)
maxratio = 0.3
vsearch_fastq_qmax = 93     # vsearch --fastq_qmax

# region -> (left-boundary prefix, right-boundary prefix) of the profile names (SeqSample.py:388-397)
REGION_PREFIXES = {"ITS2": ("3_", "4_"), "ITS1": ("1_", "2_"), "ALL": ("1_", "4_")}
