"""File formats of the reference's logger, so that its plotting tools read GPU-trained runs unchanged.

Mirrors /root/reference/src/rl/windows_workspace/spinup/utils/logx.py: ``Logger`` (progress.txt: tab-separated, header
row first, :60-75,236-257; config.json, :76-103) and the column conventions of ``EpochLogger.log_tabular``
(:66-96: ``Average<key>``, ``Std<key>``, ``Max<key>``, ``Min<key>``, or the bare key with ``average_only``).
Statistics come in already reduced over environments and ranks (ml4ca_stats5 / ml4ca_episode_stats + all-reduce), so
there is no per-sample ``store``.
"""
import json
import os
import os.path as osp
import time


def statistics_from5(s5):
    """[sum, sum of squares, count, min, max] -> (mean, std, min, max) like mpi_statistics_scalar (population std)."""
    s, q, c, lo, hi = [float(v) for v in s5]
    if c <= 0:
        return float("nan"), float("nan"), float("nan"), float("nan")
    mean = s / c
    return mean, max(q / c - mean * mean, 0.0) ** 0.5, lo, hi


class Logger(object):
    def __init__(self, output_dir=None, output_fname='progress.txt', exp_name=None, rank=0):
        self.rank = rank
        self.output_dir = output_dir or osp.join(os.getcwd(), "experiments", "%i" % int(time.time()))
        self.output_file = None
        if rank == 0:
            os.makedirs(self.output_dir, exist_ok=True)
            self.output_file = open(osp.join(self.output_dir, output_fname), 'w')
        self.first_row = True
        self.log_headers = []
        self.log_current_row = {}
        self.exp_name = exp_name

    def log_tabular(self, key, val):
        if self.first_row:
            self.log_headers.append(key)
        else:
            assert key in self.log_headers, "Trying to introduce a new key %s that you didn't include in the first iteration" % key
        assert key not in self.log_current_row, "You already set %s this iteration. Maybe you forgot to call dump_tabular()" % key
        self.log_current_row[key] = val

    def log_stats(self, key, s5, with_min_and_max=False, average_only=False):
        """EpochLogger.log_tabular(key, with_min_and_max / average_only) for already-reduced statistics."""
        mean, std, lo, hi = statistics_from5(s5)
        self.log_tabular(key if average_only else 'Average' + key, mean)
        if not average_only:
            self.log_tabular('Std' + key, std)
        if with_min_and_max:
            self.log_tabular('Max' + key, hi)
            self.log_tabular('Min' + key, lo)

    def save_config(self, config):
        out = {k: (v if isinstance(v, (int, float, str, bool, list, dict, type(None))) else str(v)) for k, v in config.items()}
        if self.exp_name is not None:
            out['exp_name'] = self.exp_name
        if self.rank == 0:
            with open(osp.join(self.output_dir, "config.json"), 'w') as fh:
                fh.write(json.dumps(out, separators=(',', ':\t'), indent=4, sort_keys=True))

    def dump_tabular(self, quiet=True):
        vals = [self.log_current_row.get(key, "") for key in self.log_headers]
        if self.rank == 0:
            if not quiet:
                w = max(15, max(len(k) for k in self.log_headers))
                print("-" * (22 + w))
                for k, v in zip(self.log_headers, vals):
                    print(("| %" + "%d" % w + "s | %15s |") % (k, "%8.3g" % v if hasattr(v, "__float__") else v))
                print("-" * (22 + w), flush=True)
            if self.first_row:
                self.output_file.write("\t".join(self.log_headers) + "\n")
            self.output_file.write("\t".join(map(str, vals)) + "\n")
            self.output_file.flush()
        self.log_current_row.clear()
        self.first_row = False
