from .CVRPEnv import CVRPEnv, Reset_State, Step_State
from .CVRPModel import CVRPModel
from .utils import augment_xy_data_by_8_fold, check_feasible, rollout, seed_everything
