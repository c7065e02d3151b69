from .TSPEnv import TSPEnv, Reset_State, Step_State
from .TSPModel import TSPModel
from .utils import augment_xy_data_by_8_fold, check_feasible, rollout, seed_everything
