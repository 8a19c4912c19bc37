"""Drop-in for the reference's MMBertForPretraining module (/root/reference/MMBertForPretraining.py), so that
``from MMBertForPretraining import MMBertForPretraining`` in train.py / sampling.py resolves to the B200-native
implementation."""
from msa_b200.api import (MMBertForPretraining, MMBertModel, MMBertPreTrainingHeads,  # noqa: F401
                          JointEmbeddings, CPC)
