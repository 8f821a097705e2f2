class Environment:  # placeholder, replaced below
    pass
