"""Running averages with the reference's interface (/root/reference/mano_train/evaluation/evalutils.py:1-30):
``AverageMeters.add_loss_value(name, value, n)`` and ``.average_meters[name].avg / .val / .sum / .count`` are what
``traineval.py`` / ``Monitor`` read after an epoch."""


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class AverageMeters(object):
    def __init__(self):
        self.average_meters = {}

    def add_loss_value(self, loss_name, loss_val, n=1):
        if loss_name not in self.average_meters:
            self.average_meters[loss_name] = AverageMeter()
        self.average_meters[loss_name].update(loss_val, n=n)
