from .manenv import ManEnv
