"""Re-export of the synthetic input generator (regcn_b200/synth.py) for the oracle-side scripts."""
from regcn_b200.synth import *  # noqa: F401,F403
from regcn_b200.synth import SHAPES, answers_of, fill_state_dict, make_case, make_snapshot  # noqa: F401
