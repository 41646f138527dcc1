import logging


def get_root_logger(*args, **kwargs):
    return logging.getLogger('basicsr')


def imwrite(*args, **kwargs):          # `modules/swin.py:9` imports these two; nothing on the oracle path calls them
    raise RuntimeError('basicsr is not installed: imwrite is a stub of the oracle shims')


def tensor2img(*args, **kwargs):
    raise RuntimeError('basicsr is not installed: tensor2img is a stub of the oracle shims')
