"""print + append-to-file logger (reference: utility/logging.py:4-17)."""
import os
from datetime import datetime


class Logger:
    def __init__(self, filename, is_debug, path="./logs/"):
        self.filename, self.path, self.log_ = filename, path, not is_debug

    def logging(self, s):
        s = str(s)
        stamp = datetime.now().strftime("%Y-%m-%d %H:%M: ")
        print(stamp, s)
        if self.log_:
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, self.filename), "a+") as f:
                f.write(stamp + " " + s + "\n")
