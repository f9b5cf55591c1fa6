"""Running averages behind the reference's meter interface
(/root/reference/mano_train/evaluation/evalutils.py:1-30): ``AverageMeters.add_loss_value(name, value, n)`` feeds
``average_meters[name]``, whose ``val`` (last value), ``sum``, ``count`` and ``avg`` are what traineval.py and the
Monitor read after an epoch.  ``epoch_pass`` fills them once per pass from the device-side loss log."""


class AverageMeter(object):
    __slots__ = ("val", "sum", "count")

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.sum, self.count = 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n

    @property
    def avg(self):
        return self.sum / self.count if self.count else 0


class AverageMeters(object):
    def __init__(self):
        self.average_meters = {}

    def add_loss_value(self, loss_name, loss_val, n=1):
        self.average_meters.setdefault(loss_name, AverageMeter()).update(loss_val, n=n)
