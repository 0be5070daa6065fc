"""Run log with the reference's interface (utility/logging.py:4-17): `Logger(filename, is_debug, path).logging(text)` echoes
a minute-resolution time stamp + text to stdout and, unless `is_debug`, appends the same line to `<path>/<filename>`.

Unlike the reference the log directory is created on demand (upstream a missing ./logs/ is a crash unless --debug is set) and
the file is opened once, line-buffered, instead of once per message."""
import os
import time

STAMP = "%Y-%m-%d %H:%M: "


class Logger:
    def __init__(self, filename, is_debug, path="./logs/"):
        self.filename = filename
        self.path = path
        self.log_ = not is_debug          # attribute name kept: callers of the reference read it
        self._sink = None

    def _file(self):
        if self._sink is None:
            os.makedirs(self.path, exist_ok=True)
            self._sink = open(os.path.join(self.path, self.filename), "a+", buffering=1)
        return self._sink

    def logging(self, s):
        text = str(s)
        stamp = time.strftime(STAMP)
        print(stamp, text)
        if self.log_:
            self._file().write("%s %s\n" % (stamp, text))

    def close(self):
        if self._sink is not None:
            self._sink.close()
            self._sink = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
