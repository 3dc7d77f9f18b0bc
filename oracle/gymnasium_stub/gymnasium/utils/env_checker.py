def check_env(env, *args, **kwargs):
    return None
