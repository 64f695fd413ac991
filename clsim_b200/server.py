"""Python twin of the in-process server seam (clsim_b200/host/I3CLSimServerInProcess.{h,cxx}).

Semantics of the reference's I3CLSimServer / I3CLSimClient pair without the sockets
(private/clsim/I3CLSimServer.cxx): any number of clients submit bunches, every converter is served by
five worker threads that each do ``EnqueueSteps`` then ``GetConversionResult`` (:126-135, 310-343), a bunch
goes to whichever worker is idle first (the ROUTER/DEALER pair balances on readiness), results return
to the client that sent the bunch under the client's own identifier, in any order.  Bunch sizes are
harmonised over the converters (:95-113).  A converter error is fatal: the server is marked failed, every
waiting client is woken with the error, and nothing is made up in place of the lost result.

The converters are ``I3CLSimStepToPhotonConverterCUDA`` objects (converter.py), one per GPU: this is the
reference's multi-device topology -- ONE process feeding N devices (private/clsim/I3CLSimModule.cxx:611-638).
ctypes releases the interpreter lock for the duration of every C-ABI call, so the workers overlap.
"""
import collections
import math
import threading


class ServerFailure(RuntimeError):
    pass


class _Mailbox(object):
    def __init__(self):
        self.cond = threading.Condition()
        self.results = collections.deque()
        self.failure = None


class I3CLSimServerInProcess(object):
    QUEUE_DEPTH = 5   # worker threads per converter (I3CLSimServer.cxx:125)

    def __init__(self, converters):
        self._converters = list(converters)
        if not self._converters:
            raise RuntimeError("Need at least 1 I3CLSimStepToPhotonConverter")
        self._workgroupSize = 0
        self._maxBunchSize = 0
        for c in self._converters:
            if c is None or not c.IsInitialized():
                raise RuntimeError("All I3CLSimStepToPhotonConverters must be initialized")
            g = c.GetWorkgroupSize()
            self._workgroupSize = g if self._workgroupSize == 0 else self._workgroupSize * g // math.gcd(self._workgroupSize, g)
            if self._maxBunchSize == 0:
                self._maxBunchSize = c.GetMaxNumWorkitems()
            else:
                m = min(self._maxBunchSize, c.GetMaxNumWorkitems())
                m -= m % self._workgroupSize
                if m == 0:
                    raise RuntimeError("maximum bunch sizes are incompatible with kernel work group sizes.")
                self._maxBunchSize = m
        self._cond = threading.Condition()
        self._frontend = collections.deque()   # (steps, internal id)
        self._clients = {}                     # internal id -> (mailbox, external id)
        self._next_id = 0
        self._shutdown = False
        self._failure = None
        self._threads = []
        for index in range(len(self._converters)):
            for _ in range(self.QUEUE_DEPTH):
                t = threading.Thread(target=self._worker, args=(index,), daemon=True)
                t.start()
                self._threads.append(t)

    def GetWorkgroupSize(self):
        return self._workgroupSize

    def GetMaxNumWorkitems(self):
        return self._maxBunchSize

    def Failure(self):
        return self._failure

    def GetStatistics(self):
        out = {}
        for i, c in enumerate(self._converters):
            post = "" if len(self._converters) == 1 else "_%d" % i
            for k, v in c.GetStatistics().items():
                out[k + post] = v
        return out

    def Connect(self):
        return I3CLSimClientInProcess(self, _Mailbox())

    def Close(self):
        with self._cond:
            self._shutdown = True
            self._cond.notify_all()
        for t in self._threads:
            t.join()
        self._threads = []

    # ---- internals ---------------------------------------------------------------------------------
    def _submit(self, mailbox, steps, external_id):
        with self._cond:
            if self._shutdown:
                raise RuntimeError("I3CLSimServerInProcess is shutting down")
            if self._failure is not None:
                raise ServerFailure("I3CLSimServerInProcess: a converter failed: " + self._failure)
            internal = self._next_id
            self._next_id = (self._next_id + 1) & 0xffffffff
            if internal in self._clients:
                raise RuntimeError("Repeated client ID")
            self._clients[internal] = (mailbox, external_id)
            self._frontend.append((steps, internal))
            self._cond.notify()

    def _fail(self, what):
        with self._cond:
            if self._failure is None:
                self._failure = what or "unknown error"
            waiting = [m for m, _ in self._clients.values()]
            self._clients.clear()
            self._frontend.clear()
        for box in waiting:
            with box.cond:
                box.failure = self._failure
                box.cond.notify_all()

    def _worker(self, index):
        conv = self._converters[index]
        while True:
            with self._cond:
                while not self._shutdown and not self._frontend:
                    self._cond.wait()
                if not self._frontend:
                    return
                steps, internal = self._frontend.popleft()
            try:
                conv.EnqueueSteps(steps, internal)
                result = conv.GetConversionResult()   # the next result, not necessarily this bunch's (:318-321)
            except Exception as e:   # fatal in the reference
                self._fail(str(e))
                return
            with self._cond:
                entry = self._clients.pop(result.identifier, None)
            if entry is None:
                continue   # "Unknown client ID" (:221-224)
            box, external = entry
            result.identifier = external
            with box.cond:
                box.results.append(result)
                box.cond.notify()


class I3CLSimClientInProcess(object):
    def __init__(self, server, mailbox):
        self._server, self._mailbox = server, mailbox
        self.workgroupSize = server.GetWorkgroupSize()
        self.maxBunchSize = server.GetMaxNumWorkitems()
        self._pending = 0

    def GetWorkgroupSize(self):
        return self.workgroupSize

    def GetMaxNumWorkitems(self):
        return self.maxBunchSize

    def EnqueueSteps(self, steps, identifier):
        self._server._submit(self._mailbox, steps, identifier)
        self._pending += 1

    def GetConversionResult(self):
        if self._pending == 0:
            return None   # I3CLSimServer.cxx:394-419: nothing outstanding
        box = self._mailbox
        with box.cond:
            while not box.results and box.failure is None:
                box.cond.wait()
            if not box.results:
                raise ServerFailure("I3CLSimServerInProcess: a converter failed: " + box.failure)
            self._pending -= 1
            return box.results.popleft()
