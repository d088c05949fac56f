"""Drop-in for the reference's model/SR/MyEfficientLFNetV4_5.py (its FastConvSSM branch, the one that runs
when mamba_ssm is not installed): same module-level symbols (`get_model`, `get_loss`, `weights_init`), same
state_dict, forward on liblfsr_b200 (sm_100a) kernels. Discovered by test.py / inference.py via
importlib.import_module('model.SR.' + args.model_name) exactly like the reference (test.py:29-31)."""
import lfsr_b200  # noqa: F401  (puts the package on sys.path / checks the native library lazily)
from lfsr_b200.lfnets.my_efficient_lfnet_v4_5 import get_model, get_loss, weights_init  # noqa: F401
