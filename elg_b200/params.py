"""Parameter containers with the reference's module/attribute names, so `state_dict()` and
`load_state_dict()` use the released checkpoint keys unchanged.  The modules hold parameters
only; the forward computation lives in libelg_b200.so.

reference: CVRP/models.py:7-25,199-209,232-247,276-297,506-562; TSP/models.py:7-22,134-172,206-225,387-423
"""
import torch
import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("elg_b200 modules hold parameters only; the computation runs in the CUDA library")


class Norm(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.InstanceNorm1d(dim, affine=True, track_running_stats=False)


class FeedForward(_Holder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.W1 = nn.Linear(dim, hidden)
        self.W2 = nn.Linear(hidden, dim)


class EncoderLayer(_Holder):
    def __init__(self, problem, **mp):
        super().__init__()
        E, HD = mp["embedding_dim"], mp["head_num"] * mp["qkv_dim"]
        self.Wq = nn.Linear(E, HD, bias=False)
        self.Wk = nn.Linear(E, HD, bias=False)
        self.Wv = nn.Linear(E, HD, bias=False)
        self.multi_head_combine = nn.Linear(HD, E)
        if problem == "cvrp":
            self.add_n_normalization_1 = Norm(E)
            self.feed_forward = FeedForward(E, mp["ff_hidden_dim"])
            self.add_n_normalization_2 = Norm(E)
        else:
            self.addAndNormalization1 = Norm(E)
            self.feedForward = FeedForward(E, mp["ff_hidden_dim"])
            self.addAndNormalization2 = Norm(E)


class Encoder(_Holder):
    def __init__(self, problem, **mp):
        super().__init__()
        self.model_params = mp
        E = mp["embedding_dim"]
        if problem == "cvrp":
            self.embedding_depot = nn.Linear(2, E)
            self.embedding_node = nn.Linear(3, E)
        else:
            self.embedding = nn.Linear(2, E)
        self.layers = nn.ModuleList([EncoderLayer(problem, **mp) for _ in range(mp["encoder_layer_num"])])


class LocalPolicy(_Holder):
    """local_policy_att parameters (CVRP/models.py:7-25)."""

    def __init__(self, problem, mp, idx=0):
        super().__init__()
        e, hd = mp["local_att_hidden_dim"], mp["local_att_head_num"] * mp["local_att_qkv_dim"]
        self.local_size = mp["local_size"][idx]
        self.init_emb = nn.Linear(3 if (problem == "cvrp" and mp.get("demand", True)) else 2, e)
        self.cur_token_emb = nn.Parameter(torch.Tensor(e))
        self.cur_token_emb.data.uniform_(-1, 1)
        self.Wq = nn.Linear(e, hd, bias=False)
        self.Wk = nn.Linear(e, hd, bias=False)
        self.Wv = nn.Linear(e, hd, bias=False)
        self.multi_head_combine = nn.Linear(hd, e)


class Decoder(_Holder):
    def __init__(self, problem, **mp):
        super().__init__()
        self.model_params = mp
        self._problem = problem
        E, HD = mp["embedding_dim"], mp["head_num"] * mp["qkv_dim"]
        if problem == "cvrp":
            self.Wq_last = nn.Linear(E + 1, HD, bias=False)
        else:
            self.Wq_first = nn.Linear(E, HD, bias=False)
            self.Wq_last = nn.Linear(E, HD, bias=False)
        self.Wk = nn.Linear(E, HD, bias=False)
        self.Wv = nn.Linear(E, HD, bias=False)
        self.multi_head_combine = nn.Linear(HD, E)
        self.local = False

    def add_local_policy(self, device, idx=0):
        """Must be called before load_state_dict, as in the reference (CVRP/test.py:76-78)."""
        if self._problem == "cvrp":
            self.local_policies = nn.ModuleList(
                [LocalPolicy("cvrp", self.model_params, i).to(device) for i in range(self.model_params["ensemble_size"])])
        else:
            self.local_policy_0 = LocalPolicy("tsp", self.model_params).to(device)
        self.local = True
