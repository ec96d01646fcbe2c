from . import rank_zero  # noqa: F401


class CombinedLoader:
    def __init__(self, *a, **k):
        pass


def move_data_to_device(batch, device):
    return batch
