"""Epoch logger with the reference's module-level API
(tonic/utils/logger.py: initialize / store / dump / show_progress / get_path /
log / warning / error; `log.csv`, `config.yaml`, `script.py` in the run
directory).  The logger itself is outside the hot path (SURVEY.md section 2,
row 13); it is kept API-compatible because the trainer and agents report
through it once per update / epoch.

Difference from the reference: instead of keeping every stored array until
`dump()`, values are folded into running aggregates (count, sum, sum of
squares, min, max), and device code can hand over pre-reduced aggregates with
`store_aggregate` so that no per-step device->host copy is needed.
"""

import datetime
import os
import time

import numpy as np
import yaml

current_logger = None


class _Aggregate:
    __slots__ = ('count', 'total', 'total_sq', 'low', 'high', 'stats', 'items', 'integral')

    def __init__(self, stats):
        self.count, self.total, self.total_sq = 0, 0.0, 0.0
        self.low, self.high = np.inf, -np.inf
        self.stats = stats
        self.items = 0          # number of stored values (reference '/size')
        self.integral = True    # every stored value was a Python / numpy integer

    def add(self, count, total, total_sq, low, high, items=1):
        self.count += count
        self.total += total
        self.total_sq += total_sq
        self.low = min(self.low, low)
        self.high = max(self.high, high)
        self.items += items

    def mean(self):
        return self.total / max(self.count, 1)

    def std(self):
        mean = self.mean()
        return float(np.sqrt(max(self.total_sq / max(self.count, 1) - mean * mean, 0.0)))


def _host(value):
    if hasattr(value, 'detach'):
        value = value.detach().cpu().numpy()
    return np.asarray(value, np.float64)


class Logger:
    def __init__(self, path=None, width=60, script_path=None, config=None):
        self.path = path or str(time.time())
        self.log_file_path = os.path.join(self.path, 'log.csv')
        self.width = width
        if script_path:
            with open(script_path) as source:
                script = source.read()
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, 'script.py'), 'w') as target:
                target.write(script)
        if config:
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, 'config.yaml'), 'w') as target:
                yaml.dump(config, target)
        self.columns = []
        self.epoch = {}
        self.start_time = time.time()
        self.last_progress = None
        self.rows_written = 0

    # -- collecting -----------------------------------------------------------
    def store(self, key, value, stats=False, items=1):
        """`items`: how many values of the reference's per-call list this call stands for
        (the reference stores one value per call and reports the list length as '/size')."""
        integral = isinstance(value, (int, np.integer)) and not isinstance(value, bool)
        v = _host(value).ravel()
        if v.size == 0:
            return
        self.store_aggregate(key, v.size, v.sum(), np.square(v).sum(), v.min(), v.max(),
                             stats=stats, items=items, integral=integral)

    def store_aggregate(self, key, count, total, total_sq, low, high, stats=False, items=1,
                        integral=False):
        agg = self.epoch.get(key)
        if agg is None:
            agg = self.epoch[key] = _Aggregate(stats)
        agg.add(count, float(total), float(total_sq), float(low), float(high), items)
        agg.integral = agg.integral and integral

    # -- reporting --------------------------------------------------------------
    def _row(self):
        row = {}
        for key, agg in self.epoch.items():
            if agg.stats:
                row[key + '/mean'] = agg.mean()
                row[key + '/std'] = agg.std()
                row[key + '/min'] = agg.low
                row[key + '/max'] = agg.high
                row[key + '/size'] = agg.items
            else:
                row[key] = agg.mean()
                if agg.integral and row[key] == int(row[key]):    # train/steps, episodes, epochs
                    row[key] = int(row[key])
        return row

    def dump(self):
        row = self._row()
        new = [k for k in row if k not in self.columns]
        if new and self.columns:
            warning(f'Logging new keys {new}')
        old_columns = list(self.columns)
        self.columns = sorted(set(self.columns) | set(new))
        self._print(row)
        os.makedirs(self.path, exist_ok=True)
        line = ','.join(str(row.get(k)) for k in self.columns)
        if self.rows_written == 0:
            log(f'Logging data to {self.log_file_path}')
            with open(self.log_file_path, 'w') as f:
                f.write(','.join(self.columns) + '\n' + line + '\n')
        elif new:       # re-write earlier rows with 'None' in the new columns
            with open(self.log_file_path) as f:
                lines = f.read().splitlines()[1:]
            with open(self.log_file_path, 'w') as f:
                f.write(','.join(self.columns) + '\n')
                for text in lines:
                    cells = dict(zip(old_columns, text.split(',')))
                    f.write(','.join(cells.get(k, 'None') for k in self.columns) + '\n')
                f.write(line + '\n')
        else:
            with open(self.log_file_path, 'a') as f:
                f.write(line + '\n')
        self.rows_written += 1
        self.epoch.clear()
        self.last_progress = None
        return row

    def _print(self, row):
        print()
        shown = set()
        for key in sorted(row):
            *groups, leaf = key.split('/')
            for depth in range(len(groups)):
                prefix = '/'.join(groups[:depth + 1])
                if prefix not in shown:
                    shown.add(prefix)
                    print('  ' * depth + groups[depth].replace('_', ' '))
            value = row[key]
            text = f'{value:,}' if isinstance(value, (int, np.integer)) else f'{value:8.3g}'
            left = '  ' * len(groups) + leaf.replace('_', ' ')
            print(left + ' ' * max(1, self.width - len(left) - len(text)) + text)
        print()

    def show_progress(self, steps, num_epoch_steps, num_steps):
        epoch_steps = (steps - 1) % num_epoch_steps + 1
        progress = int(self.width * epoch_steps / num_epoch_steps)
        if progress == self.last_progress:
            return
        per_step = (time.time() - self.start_time) / max(steps, 1)
        left_epoch = datetime.timedelta(seconds=int((num_epoch_steps - epoch_steps) * per_step))
        left_total = datetime.timedelta(seconds=int(max(num_steps - steps, 0) * per_step))
        msg = f'Time left:  epoch {left_epoch}  total {left_total}'.center(self.width)
        print('\r' + '#' * 0 + msg, end='')
        self.last_progress = progress


def initialize(*args, **kwargs):
    global current_logger
    current_logger = Logger(*args, **kwargs)
    return current_logger


def get_current_logger():
    global current_logger
    if current_logger is None:
        current_logger = Logger()
    return current_logger


def store(*args, **kwargs):
    return get_current_logger().store(*args, **kwargs)


def store_aggregate(*args, **kwargs):
    return get_current_logger().store_aggregate(*args, **kwargs)


def dump(*args, **kwargs):
    return get_current_logger().dump(*args, **kwargs)


def show_progress(*args, **kwargs):
    return get_current_logger().show_progress(*args, **kwargs)


def get_path():
    return get_current_logger().path


def log(msg, color='green'):
    print(msg)


def warning(msg, color='yellow'):
    print('Warning: ' + msg)


def error(msg, color='red'):
    print('Error: ' + msg)
