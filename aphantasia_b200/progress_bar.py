"""Console progress read-out with the reference's API (ProgressBar(n).upd(); /root/reference/aphantasia/progress_bar.py:53-112).
Prints the same `rate Xs` per-step figure the reference uses as its only speed read-out (progress_bar.py:98)."""
import sys
import time


def shortime(sec):
    if sec < 60: return '%d' % sec
    if sec < 3600: return '%d:%02d' % ((sec / 60) % 60, sec % 60)
    if sec < 86400: return '%d:%02d:%02d' % (sec / 3600, (sec / 60) % 60, sec % 60)
    return '%dd %d:%02d:%02d' % (sec / 86400, (sec / 3600) % 24, (sec / 60) % 60, sec % 60)


class ProgressBar(object):
    def __init__(self, task_num=0, bar_width=50, start=True):
        self.task_num, self.bar_width, self.completed = task_num, bar_width, 0
        if start: self.start()

    def start(self, task_num=None):
        if task_num is not None: self.task_num = task_num
        self.start_time = time.time()

    def upd(self, msg=None):
        self.completed += 1
        elapsed = time.time() - self.start_time + 1e-13
        rate = elapsed / self.completed
        if self.task_num > 0:
            frac = self.completed / float(self.task_num)
            eta = int(elapsed * (1 - frac) / frac + 0.5)
            line = '\r[%s] %d/%d, rate %.3gs, time %ss, left %ss %s' % (
                'X' * int(self.bar_width * frac) + '-' * (self.bar_width - int(self.bar_width * frac)), self.completed, self.task_num,
                rate, shortime(elapsed), shortime(eta), '' if msg is None else str(msg))
        else:
            line = '\rcompleted %d, time %ds, %.1f steps/s' % (self.completed, int(elapsed + 0.5), 1. / rate)
        sys.stdout.write(line)
        if self.completed == self.task_num:
            sys.stdout.write('\n')
            # last preview frame submitted: wait for the asynchronous JPEG encoder before the caller reads the output
            # directory (clip_fft.py:311-313 runs ffmpeg and img_list right after the loop)
            from .utils import _drain_saves
            _drain_saves()
        sys.stdout.flush()
        return self.completed

    def reset(self, count=None, newline=False):
        self.start_time = time.time()
        if count is not None: self.task_num = count
        if newline: sys.stdout.write('\n\n')


ProgressIPy = ProgressBar
