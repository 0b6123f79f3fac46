"""Wall-clock report, same behaviour as ``co_monitor.geometry.decorators.timer``
(reference: co_monitor/geometry/decorators.py:6-18): the wrapped callable's result is returned
unchanged and ``Execution finished in  x.xxs`` is printed.  The reference applies it to the
``Simulation`` *class* (simulation.py:30), so ``Simulation`` is a plain function there; same here."""
import functools
import time


def timer(function):
    @functools.wraps(function)
    def wrapper(*args, **kwargs):
        t0 = time.time()
        value = function(*args, **kwargs)
        print(f"\nExecution finished in {time.time() - t0: .2f}s")
        return value

    return wrapper
