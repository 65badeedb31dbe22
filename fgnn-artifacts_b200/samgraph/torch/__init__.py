from samgraph.torch.adapter import *  # noqa: F401,F403
