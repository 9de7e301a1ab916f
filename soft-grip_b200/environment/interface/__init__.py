from .environment import Env
