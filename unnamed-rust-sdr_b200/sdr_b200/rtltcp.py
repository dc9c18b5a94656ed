"""rtl_tcp wire ingest -- host-side mirror of src/rtltcp.rs:7-168.

    RtlTcp (builder, :7-77)            -> RtlTcp(...).address().rate().frequency().gain().rtlagc().listen()
    RtlTcpConnection::connect (:96-109) -> 12-byte greeting read, then SetSampleRate
    RtlTcpConnection::command (:111-134) -> 1 command byte + u32 big-endian argument; the sample-rate check of
                                           :123-131 (the reference panics; here ValueError)
    RtlTcpSignal::next (:158-164)      -> RtlTcpSignal.next_raw(n): raw I/Q bytes in blocks (the GPU unpacks them);
                                           any read error ends the stream (:159 `.ok()?`)

No sample arithmetic happens here: the bytes go to the device as they arrive (ops.Fir / FftPlan / FmStereo take u8 IQ).
"""
import socket
import struct

import numpy as np

from . import ops
from .signal import Signal

CMD_SET_FREQUENCY = 0x01        # rtltcp.rs:113
CMD_SET_SAMPLE_RATE = 0x02      # :114
CMD_SET_TUNER_GAIN_MODE = 0x03  # :115
CMD_SET_TUNER_GAIN = 0x04       # :116
CMD_SET_RTL_AGC = 0x08          # :117


def command_bytes(cmd, arg):
    """write_u8(cmd); write_u32::<BigEndian>(arg)   (rtltcp.rs:120-121)"""
    return struct.pack(">BI", cmd, int(arg) & 0xFFFFFFFF)


def check_sample_rate(rate):
    """the ranges of rtltcp.rs:123-131 (the reference panics outside them)"""
    if not (225001 <= rate <= 300000) and not (900001 <= rate <= 3200000):
        raise ValueError("bad sample rate for rtltcp: %r" % (rate,))


def gain_tenths_db(gain):
    """rtltcp.rs:64-68: (gain * 10.0).round() as u32 for gain > 0, else 0 (f32 arithmetic, round half away)"""
    g = np.float32(gain)
    if not g > 0:
        return 0
    v = np.float32(g * np.float32(10.0))
    return int(np.floor(np.float64(v) + 0.5))


class RtlTcpConnection:
    def __init__(self, rate, addr, timeout=None):
        self.sock = socket.create_connection(addr, timeout=timeout)
        self.id = self._read_exact(12)  # rtltcp.rs:100-101
        if len(self.id) != 12:
            self.sock.close()
            raise ConnectionError("rtl_tcp greeting: got %d of 12 bytes" % len(self.id))
        self.rate = rate
        self.command(CMD_SET_SAMPLE_RATE, rate)  # :107

    def _read_exact(self, n):
        parts, got = [], 0
        while got < n:
            try:
                b = self.sock.recv(min(n - got, 1 << 20))
            except OSError:
                break
            if not b:
                break
            parts.append(b)
            got += len(b)
        return b"".join(parts)

    def command(self, cmd, arg):
        self.sock.sendall(command_bytes(cmd, arg))
        if cmd == CMD_SET_SAMPLE_RATE:
            check_sample_rate(arg)
            self.rate = arg

    def read_bytes(self, n_samples):
        """up to n_samples I/Q pairs; a short (even-length) block means the stream ended"""
        b = self._read_exact(2 * n_samples)
        return np.frombuffer(b[:len(b) // 2 * 2], np.uint8)

    def listen(self):
        return RtlTcpSignal(self)

    def close(self):
        try:
            self.sock.close()
        except OSError:
            pass


class RtlTcpSignal(Signal):
    """Signal<Sample = Complex<f32>> at the connection's rate (rtltcp.rs:151-168); consumers that understand u8 IQ
    pull `next_raw` and unpack on the GPU."""
    dtype = np.complex64

    def __init__(self, conn):
        self.conn = conn
        self._rate = np.float32(conn.rate)
        self.done = False

    def rate(self):
        return self._rate

    def next_raw(self, n):
        if self.done:
            return np.empty(0, np.uint8)
        b = self.conn.read_bytes(n)
        if len(b) < 2 * n:
            self.done = True
        return b

    def next_block(self, n):
        raw = self.next_raw(n)
        if len(raw) == 0:
            return np.empty(0, np.complex64)
        return ops.unpack_u8iq(raw)


class RtlTcp:
    """builder of src/rtltcp.rs:7-58 with the same defaults (:17-27)"""

    def __init__(self):
        self._addr = ("127.0.0.1", 1234)
        self._rate = 1800000
        self._frequency = 100000000
        self._gain = None
        self._rtlagc = False

    def address(self, addr):
        if isinstance(addr, str):
            host, _, port = addr.rpartition(":")
            addr = (host, int(port))
        self._addr = addr
        return self

    def rate(self, rate):
        self._rate = int(rate)
        return self

    def frequency(self, frequency):
        self._frequency = int(frequency)
        return self

    def gain(self, gain):
        self._gain = gain
        return self

    def rtlagc(self, rtlagc):
        self._rtlagc = bool(rtlagc)
        return self

    def listen(self, timeout=None):
        """rtltcp.rs:60-77: connect (greeting + SetSampleRate), SetFrequency, gain mode (+ gain), SetRtlAgc"""
        conn = RtlTcpConnection(self._rate, self._addr, timeout=timeout)
        conn.command(CMD_SET_FREQUENCY, self._frequency)
        if self._gain is not None:
            conn.command(CMD_SET_TUNER_GAIN_MODE, 1)
            conn.command(CMD_SET_TUNER_GAIN, gain_tenths_db(self._gain))
        else:
            conn.command(CMD_SET_TUNER_GAIN_MODE, 0)
        conn.command(CMD_SET_RTL_AGC, 1 if self._rtlagc else 0)
        return conn.listen()
