"""Run a worker module on one dedicated daemon thread (the Blender add-on reaches the renderer this way:
`worker = OnDemandProxy(lambda: DaemonModule(lambda: importlib.import_module('ptina_b200.worker')))`, blender.py:13-33).

Same contract as the reference's ptina/tools/mtworker.py:12-89: every attribute of the wrapped module that is callable runs
on the daemon thread while the caller blocks until it returns; exceptions raised there are printed and turn into a `None`
result; non-callables pass through; `direct_launch(func)` runs an arbitrary function on the thread.  The CUDA context of
ptina_b200 is created by whatever thread first calls `worker.init()`, which this shim makes the daemon thread -- so all
device work of a process stays on one thread, as the library (single-threaded, non-re-entrant) requires."""
import functools
import queue
import threading
import traceback


class DaemonWorker:
    def __init__(self):
        self._jobs = queue.Queue()
        self._thread = threading.Thread(target=self._main, daemon=True)
        self._thread.start()

    def _main(self):
        while True:
            job = self._jobs.get()
            try:
                job['ret'] = job['func']()
            except BaseException:
                print(traceback.format_exc())
                job['ret'] = None
            job['done'].set()

    def launch(self, func):
        """Run `func()` on the daemon thread, block until it is done, return its result (None if it raised)."""
        if threading.current_thread() is self._thread:        # re-entrant call from the worker thread itself
            return func()
        job = {'func': func, 'ret': None, 'done': threading.Event()}
        self._jobs.put(job)
        job['done'].wait()
        return job['ret']


class DaemonModule:
    def __init__(self, getmodule):
        self._worker = DaemonWorker()
        self._module = self._worker.launch(getmodule)

    def direct_launch(self, func):
        return self._worker.launch(func)

    def _wrap(self, obj):
        if not callable(obj):
            return obj

        @functools.wraps(obj)
        def wrapped(*args, **kwargs):
            return self._wrap(self._worker.launch(lambda: obj(*args, **kwargs)))
        return wrapped

    def __getattr__(self, name):
        return self._wrap(getattr(self._module, name))


class OnDemandProxy:
    """Import / construct the module on first use (double-checked under a lock)."""
    def __init__(self, getmodule):
        self._getmodule = getmodule
        self._module = None
        self._lock = threading.Lock()

    def __getattr__(self, name):
        if self._module is None:
            with self._lock:
                if self._module is None:
                    self._module = self._getmodule()
        return getattr(self._module, name)
