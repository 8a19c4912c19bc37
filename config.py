"""Drop-in for the reference's config.py (/root/reference/config.py:1-17): same names, same values."""
import torch

DEVICE = torch.device("cuda")

total_vocab_size = 30522

# modality dimensions
TEXTDIM = 1024
MOSEIVISUALDIM = 35
MOSIVISUALDIM = 47
FUNNYVISUALDIM = 371
CMUSPEECHDIM = 74
FUNNYSPEECHDIM = 81
