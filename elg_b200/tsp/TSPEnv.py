"""Drop-in for the reference's TSPEnv (TSP/TSPEnv.py:24-184) on bit-mask device state."""
import torch

from .. import engine


class Reset_State:
    def __init__(self, problems):
        self.problems = problems
        # shape: (batch, problem, 2)


class Step_State:
    def __init__(self, BATCH_IDX, POMO_IDX):
        self.BATCH_IDX = BATCH_IDX
        self.POMO_IDX = POMO_IDX
        self.current_node = None
        self.ninf_mask = None
        self._mask_bits = None


class TSPEnv:
    _elg_fused = True

    def __init__(self, multi_width, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise engine._lib.ElgError("elg_b200.TSPEnv needs a CUDA device; there is no CPU path")
        self.problem_size = None
        self.pomo_size = multi_width
        self.tsplib = False
        self.batch_size = None
        self.BATCH_IDX = None
        self.POMO_IDX = None
        self.problems = None
        self.unscaled_problems = None
        self._dist = None
        self.selected_count = None
        self.current_node = None
        self._actions, self._solutions = [], None

    @property
    def multi_width(self):
        return self.pomo_size

    @property
    def dist(self):
        if self._dist is None and self.problems is not None:
            self._dist = engine.pairwise_dist(self.problems)
        return self._dist

    def _set_problems(self, problems, aug_factor):
        if aug_factor not in (1, 8):
            raise NotImplementedError
        problems = problems.to(self.device)
        self.problem_size = problems.size(1)
        self.problems, _ = engine.load_problems("tsp", problems, aug=aug_factor)
        self.batch_size = self.problems.size(0)
        self._dist = None
        self.BATCH_IDX = torch.arange(self.batch_size, device=self.device)[:, None].expand(self.batch_size, self.pomo_size)
        self.POMO_IDX = torch.arange(self.pomo_size, device=self.device)[None, :].expand(self.batch_size, self.pomo_size)

    def load_random_problems(self, problems, aug_factor=1):
        """problems (n, N, 2) -- TSP/TSPEnv.py:53-67."""
        self.tsplib = False
        self._set_problems(problems, aug_factor)

    def load_tsplib_problem(self, problems, unscaled_problems, aug_factor=1):
        """scaled problems (1, N, 2) + unscaled (1, N, 2) shared by all augmentations -- TSP/TSPEnv.py:69-85."""
        self._set_problems(problems, aug_factor)
        self.tsplib = True
        self.unscaled_problems = unscaled_problems
        self._unscaled_dev = unscaled_problems.to(self.device).float().expand(self.batch_size, -1, -1).contiguous()

    def reset(self):
        B, M, N = self.batch_size, self.pomo_size, self.problem_size
        self.selected_count = 0
        self.current_node = None
        self._actions, self._solutions = [], None
        self.step_state = Step_State(BATCH_IDX=self.BATCH_IDX, POMO_IDX=self.POMO_IDX)
        self._visited_bits = torch.zeros((B, M, engine.mask_words(N)), dtype=torch.int32, device=self.device)
        self._mask_bits = torch.zeros((B, M, engine.mask_words(N)), dtype=torch.int32, device=self.device)
        self._ninf_shape = (B, M, N)
        self._ninf = None
        self.step_state._mask_bits = self._mask_bits
        return Reset_State(self.problems), None, False

    def pre_step(self):
        if self._ninf is None:
            self._ninf = torch.zeros(self._ninf_shape, device=self.device)
        self.step_state.ninf_mask = self._ninf
        return self.step_state, None, False

    def step(self, selected):
        """TSP/TSPEnv.py:108-133."""
        if self._ninf is None:
            self._ninf = torch.zeros(self._ninf_shape, device=self.device)
        self.selected_count += 1
        self.current_node = selected
        self._actions.append(selected)
        engine.env_step("tsp", None, selected.to(torch.int32).contiguous(), None, self._visited_bits, self._mask_bits,
                        None, self._ninf, None)
        self.step_state.current_node = self.current_node
        self.step_state.ninf_mask = self._ninf
        self.step_state._mask_bits = self._mask_bits
        done = self.selected_count == self.problem_size
        reward = None
        if done:
            reward = self.compute_unscaled_distance() if self.tsplib else -self._get_travel_distance()
        return self.step_state, reward, done

    @property
    def selected_node_list(self):
        if self._solutions is not None:
            return self._solutions
        if not self._actions:
            return torch.zeros((self.batch_size, self.pomo_size, 0), dtype=torch.long, device=self.device)
        return torch.stack(self._actions, dim=2)

    def _finish_fused(self, solutions):
        self._solutions = solutions
        self.selected_count = solutions.shape[2]
        self.current_node = solutions[:, :, -1]

    def get_local_feature(self):
        """(cur_dist, cur_theta, relative_xy) -- TSP/TSPEnv.py:135-156."""
        if self.current_node is None:
            return None, None, None
        d, th, rel, _ = engine.cur_feature(self.problems, self.current_node)
        return d, th, rel

    def _get_travel_distance(self):
        return engine.tour_length(self.problems, self.selected_node_list)

    def compute_unscaled_distance(self, solutions=None):
        if solutions is None:
            solutions = self.selected_node_list
        return -engine.tour_length(self._unscaled_dev, solutions, rounding=True)
