"""`ValueStats`, `PriorStats`, `CombinedStats` with the reference's accounting (oinkoink/neural/stats.py:4-143):
running totals over mini-batches; the same field names, `to_dict()` keys and printed form, so a results table built
from them (`TrainingLoop._evaluate`, neural/training.py:156-171) looks the same."""
import numpy as np

CATEGORIES = (0.0, 0.5, 1.0)


def categorise_predictions(preds):
    """value in [0,1) -> {0, 0.5, 1}: thirds of the unit interval (stats.py:66-70)"""
    return np.floor(np.asarray(preds) * 3.0) / 2.0


class ValueStats():
    def __init__(self):
        self.n = 0
        self.average_value = 0.0
        self.total_loss = 0.0
        self.smallest = 1.0
        self.largest = 0.0
        self.correct = {c: 0 for c in CATEGORIES}
        self.total = {c: 0 for c in CATEGORIES}

    @property
    def loss(self):
        return float(self.total_loss) / self.n

    @property
    def accuracy(self):
        return float(sum(self.correct.values())) / self.n

    @property
    def average(self):
        return self.average_value / self.n

    def update(self, outputs, values, loss):
        outputs, values = np.asarray(outputs), np.asarray(values)
        self.n += len(values)
        self.average_value += np.sum(outputs)
        self.total_loss += loss * len(values)
        self.smallest = min(self.smallest, np.min(outputs).item())
        self.largest = max(self.largest, np.max(outputs).item())
        hit = categorise_predictions(outputs) == values
        for c in CATEGORIES:
            members = values == c
            self.total[c] += int(np.count_nonzero(members))
            self.correct[c] += int(np.count_nonzero(hit & members))

    def to_dict(self):
        d = {'Average loss': self.loss, 'Accuracy': self.accuracy, 'Smallest': self.smallest,
             'Largest': self.largest, 'Average': self.average}
        d['correct'] = {c: (self.total[c], self.correct[c]) for c in self.correct}
        return d

    def __repr__(self):
        head = "Average loss:  {:.5f}  Accuracy:  {:.5f}  Smallest:  {:.5f}  Largest:  {:.5f}  Average:  {:.5f}".format(
            self.loss, self.accuracy, self.smallest, self.largest, self.average)
        cats = "".join("  ({}, {}, {})".format(c, self.total[c], self.correct[c]) for c in self.correct)
        return head + "\nCategory, # Members, # Correct Predictions:" + cats

    categorise_predictions = staticmethod(categorise_predictions)


class PriorStats():
    def __init__(self):
        self.n = 0
        self.total_loss = 0.0
        self.correct = 0

    @property
    def loss(self):
        return float(self.total_loss) / self.n

    @property
    def accuracy(self):
        return float(self.correct) / self.n

    def update(self, outputs, values, loss):
        """a prediction is correct when the network's best move is ONE OF the label's best moves (stats.py:101-113)"""
        outputs, values = np.asarray(outputs), np.asarray(values)
        self.n += len(values)
        self.total_loss += loss * len(values)
        best = np.argmax(outputs, axis=1)
        rows = np.arange(len(values))
        self.correct += int(np.count_nonzero(values[rows, best] == np.amax(values, axis=1)))

    def to_dict(self):
        return {'Average loss': self.loss, 'Accuracy': self.accuracy}

    def __repr__(self):
        return "Average loss:  {:.5f}  Accuracy:  {:.5f}".format(self.loss, self.accuracy)


class CombinedStats:
    def __init__(self):
        self.value_stats = ValueStats()
        self.prior_stats = PriorStats()

    @property
    def loss(self):
        return self.value_stats.loss + self.prior_stats.loss

    def update(self, value_outputs, values, value_loss, prior_outputs, priors, prior_loss):
        self.value_stats.update(value_outputs, values, value_loss)
        self.prior_stats.update(prior_outputs, priors, prior_loss)

    def to_dict(self):
        d = {'prior ' + k: v for k, v in self.prior_stats.to_dict().items()}
        d.update(self.value_stats.to_dict())
        return d

    def __repr__(self):
        return "{}\n{}".format(repr(self.value_stats), repr(self.prior_stats))
