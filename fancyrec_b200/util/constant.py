"""Mirror of the reference's util/constant.py:5-6: one global device."""
import torch

device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
