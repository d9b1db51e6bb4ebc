"""One launch of each recurrence kernel shape of the bench step (MOSI, batch 2048) for an ncu capture: the 6-cell
forward / backward launch and the three decoder cells, forward then backward (8 launches)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("LSTM_PROF_ONCE", "1")
import lstm_prof  # noqa: F401,E402
