"""Mirror of the reference's util/constant.py:5-6: dataset root and one global device."""
import os

import torch

ROOT_PATH = os.environ.get("FANCYREC_ROOT_PATH", "/home/u190110105/insCar")      # util/constant.py:5 (the authors' path)
device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
