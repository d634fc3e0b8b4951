"""Re-export of the synthetic weight / input generator (it lives in the package: dreamer_b200/synthetic.py)."""
from dreamer_b200.synthetic import *  # noqa: F401,F403
from dreamer_b200.synthetic import REF_CONFIG, make_state_dict, rollout_inputs, sequence_inputs, small_config, state_dict_shapes  # noqa: F401
