def register(*args, **kwargs):
    """no-op (pearl/user_envs/__init__.py calls this at import time)"""
    return None
