"""Host-buffer front end of the sweep engine: payload bits in (pinned) HOST memory in, bit-error counts in
pinned host memory out, one batch of B streams per step, with the H2D copy of batch i + 1 overlapping the
kernels of batch i.  This is what a GUI-style caller does around the reference (`SIMO/gui/main_window.py:176-231`:
bits held by the application, one `simulate_simo` call per payload) turned into a stream of batches.

    pipe = engine.stream_host_batches(chan, num_rx, snr_lin_rows, B, S)
    for errors, tag in pipe.run((bits_i, stream_id0_i, tag_i) for ...):   # errors: int64 [B] on the host
        ...
"""
import os

import torch


def gpu_numa_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off, from sysfs (None if the platform does not say)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:                       # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f'/sys/bus/pci/devices/{bus}/numa_node').read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa(device_index):
    """Pins the calling process to the CPUs next to its GPU, so that the pinned staging buffers it allocates
    afterwards (first touch) and its copy-issuing thread live on that NUMA node.  Best effort: returns the CPU
    list it bound to, or None when the topology is unknown / the affinity call is not permitted."""
    cpus = gpu_numa_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        want = set(cpus) & allowed
        if want:
            os.sched_setaffinity(0, want)
            return sorted(want)
    except (AttributeError, OSError):
        pass
    return None


class HostBatchPipeline:
    """Double-buffered host -> device -> host pipeline around LinkEngine.simo_ber.

    Every batch is B streams of S OFDM symbols whose payload arrives as np.packbits() rows
    uint8 [B, ceil(nbits / 8)] in host memory (pinned memory makes the copy asynchronous).  Per batch, inside
    one CUDA stream order: H2D copy (copy stream) -> lte_bits_to_indices -> the SIMO-MRC chain -> D2H of the
    int64 [B] error counts into a pinned host buffer.  `depth` device input buffers let the copy of the next
    batch run while the current one computes; consecutive batches alternate between two compute streams."""

    def __init__(self, engine, chan, num_rx, snr_lin_rows, B, S, nbits=None, seed=0, noise_domain=3, fused=True,
                 depth=3, spectral=None):
        self.eng, self.chan, self.R, self.B, self.S = engine, chan, int(num_rx), int(B), int(S)
        self.nbits = int(nbits) if nbits is not None else S * engine.Nd * engine.bps
        self.nbytes = (self.nbits + 7) // 8
        self.seed, self.noise_domain, self.fused, self.depth = int(seed), int(noise_domain), bool(fused), int(depth)
        self.spectral = spectral                                # None: the engine picks (spectral link when it applies)
        dev = engine.device
        self.snr_rows = snr_lin_rows.to(dev).contiguous()
        # through the spectral link two batches in flight on two compute streams (the front of one fills the SMs the
        # tail of the other leaves idle, LinkEngine.simo_ber_batches), each with its own workspace
        nfl = engine.batches_in_flight(chan, self.B, self.R, self.S, fused=fused, spectral=spectral)
        self.wss = [engine.workspace(self.B, self.S, self.R, fading=chan.num_taps > 0, fused=fused,
                                     lazy=fused and spectral is not False) for _ in range(nfl)]
        self.ws = self.wss[0]
        self.cstreams = engine.side_streams(nfl)
        self.dev_bits = [torch.empty((self.B, self.nbytes), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.host_err = [torch.empty(self.B, dtype=torch.int64).pin_memory() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_copied = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes_per_batch = self.B * self.nbytes
        self.d2h_bytes_per_batch = self.B * 8

    def _upload(self, k, host_bits):
        main = torch.cuda.current_stream(self.eng.device)
        if host_bits.shape != (self.B, self.nbytes) or host_bits.dtype != torch.uint8:
            raise ValueError(f"a batch must be uint8 [{self.B}, {self.nbytes}] of np.packbits() rows")
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_free[k])
            self.dev_bits[k].copy_(host_bits, non_blocking=True)
            self.ev_copied[k].record(self.copy_stream)
        return main

    def _compute(self, k, stream_id0, i):
        j = i % len(self.cstreams)
        cs = self.cstreams[j]
        cs.wait_event(self.ev_copied[k])
        with torch.cuda.stream(cs):
            idx = self.eng.bits_to_indices(self.dev_bits[k], self.nbits, self.S, packed=True)
            err = self.eng.simo_ber(self.wss[j], self.chan, self.snr_rows, self.seed, stream_id0=int(stream_id0),
                                    idx=idx, nbits=self.nbits, fused=self.fused, spectral=self.spectral,
                                    noise_domain=self.noise_domain)
            self.host_err[k].copy_(err, non_blocking=True)
            self.ev_free[k].record(cs)
            self.ev_done[k].record(cs)

    def run(self, batches):
        """batches: iterable of (host_bits, stream_id0, tag).  Yields (errors, tag) per batch in order; `errors` is
        the pipeline's pinned int64 [B] buffer of that slot -- valid until `depth` more batches have been yielded."""
        main = torch.cuda.current_stream(self.eng.device)
        for k in range(self.depth):
            self.ev_free[k].record(main)
        for cs in self.cstreams:
            cs.wait_stream(main)
        it = iter(batches)
        pending = []                                            # (slot, tag) whose results are not yet handed out
        nxt = next(it, None)
        i = 0
        if nxt is not None:
            self._upload(0, nxt[0])
        while nxt is not None:
            cur, k = nxt, i % self.depth
            nxt = next(it, None)
            if nxt is not None:
                if len(pending) >= self.depth - 1:              # the slot about to be overwritten is still owed to the caller
                    ks, tag = pending.pop(0)
                    self.ev_done[ks].synchronize()
                    yield self.host_err[ks], tag
                self._upload((i + 1) % self.depth, nxt[0])
            self._compute(k, cur[1], i)
            pending.append((k, cur[2]))
            i += 1
        for cs in self.cstreams:
            main.wait_stream(cs)                                # events the caller records on its stream cover all batches
        for ks, tag in pending:
            self.ev_done[ks].synchronize()
            yield self.host_err[ks], tag
