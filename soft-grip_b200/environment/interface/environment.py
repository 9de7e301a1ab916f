class Env(object):
    """Abstract environment (mirror of ref: environment/interface/environment.py:1-10)."""

    def __init__(self, sim_start, sim_step):
        self.sim_start = sim_start
        self.sim_step = sim_step

    def step(self, *args):
        raise NotImplementedError("Not implemented")

    def reset(self):
        raise NotImplementedError("Not implemented")
