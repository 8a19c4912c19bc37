"""Drop-in for the reference's MMBertEmbedding module (/root/reference/MMBertEmbedding.py): re-exports the
B200-native classes under the reference's import path."""
from msa_b200.api import CPC, JointEmbeddings  # noqa: F401
