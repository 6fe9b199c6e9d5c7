"""i-PI socket client: lets an external MD engine (i-PI and the codes that speak its driver protocol) use an energy /
force callback of this package as its force field (reference: Interfaces/TMIPIinterface.py:7-61, class `TMIPIManger`,
the reference's spelling kept).

Wire protocol (what the reference's client reads and writes, all little-endian native types): 12-byte ASCII headers padded
with blanks; server -> client "STATUS", "POSDATA" (cell 9 f64, inverse cell 9 f64, natom i32, positions 3 natom f64, all
atomic units) and "GETFORCE"; client -> server "READY" / "HAVEDATA" and, after "GETFORCE", "FORCEREADY" + energy f64 +
natom i32 + forces 3 natom f64 (Hartree / Bohr) + virial 9 f64 + an i32 length and that many extra bytes.

Unit handling follows the reference: positions arrive in Bohr and are handed to the callback in Angstrom; the callback
returns (E [Hartree], F [J/mol/Angstrom]) like every force callback of this package and the force is sent as
F / JOULEPERHARTREE / BOHRPERA; the virial is sent as zeros.

Differences from the reference, all on the transport side (it is Python-2 code: `str` headers and `np.fromstring`):
headers are bytes, every read loops until the announced byte count has arrived (a TCP `recv` may return less), the i-PI
messages "INIT" (bead index + an init string, consumed) and "EXIT" (returns) are understood, a closed socket ends the
loop instead of raising on an empty header, and a failed connection raises instead of printing.
"""
from __future__ import annotations

import socket

import numpy as np

from ..PhysicalData import BOHRPERA, JOULEPERHARTREE

HDRLEN = 12


def _header(msg: str) -> bytes:
    return msg.encode("ascii").ljust(HDRLEN)


class TMIPIManger:
    def __init__(self, EnergyForceField=None, TCP_IP="localhost", TCP_PORT=31415, sock_=None):
        """EnergyForceField(x[N,3] Angstrom) -> (E Hartree, F[N,3] J/mol/Angstrom). `sock_` (extension): an already
        connected socket, e.g. a UNIX-domain one."""
        self.EnergyForceField = EnergyForceField
        self.hasdata = False
        self.nsteps = 0
        self.cellh = None
        self.cellih = None
        if sock_ is not None:
            self.s = sock_
        else:
            self.s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            self.s.connect((TCP_IP, TCP_PORT))

    def _recv_exact(self, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = self.s.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("i-PI server closed the socket mid-message")
            buf += chunk
        return bytes(buf)

    def _recv_header(self) -> str:
        first = self.s.recv(HDRLEN)
        if not first:
            return ""
        if len(first) < HDRLEN:
            first += self._recv_exact(HDRLEN - len(first))
        return first.decode("ascii").strip()

    def md_run(self):
        """Serves the socket until the server says EXIT or closes the connection; returns the number of force
        evaluations done."""
        energy, natom, force, vir = 0.0, 0, None, np.zeros((3, 3))
        while True:
            msg = self._recv_header()
            if msg == "" or msg == "EXIT":
                return self.nsteps
            if msg == "STATUS":
                self.s.sendall(_header("HAVEDATA" if self.hasdata else "READY"))
            elif msg == "INIT":
                self._recv_exact(4)                                          # bead index
                nbytes = int(np.frombuffer(self._recv_exact(4), np.int32)[0])
                self._recv_exact(nbytes)                                     # initialisation string: unused
            elif msg == "POSDATA":
                self.cellh = np.frombuffer(self._recv_exact(9 * 8), np.float64) / BOHRPERA
                self.cellih = np.frombuffer(self._recv_exact(9 * 8), np.float64) * BOHRPERA
                natom = int(np.frombuffer(self._recv_exact(4), np.int32)[0])
                position = (np.frombuffer(self._recv_exact(3 * natom * 8), np.float64) / BOHRPERA).reshape((-1, 3))
                energy, force = self.EnergyForceField(position)
                force = np.ascontiguousarray(np.asarray(force, np.float64).reshape(natom, 3) / JOULEPERHARTREE / BOHRPERA)
                vir = np.zeros((3, 3))
                self.hasdata = True
                self.nsteps += 1
            elif msg == "GETFORCE":
                if not self.hasdata:
                    raise Exception("GETFORCE before any POSDATA")
                extra = b"nothing"
                self.s.sendall(_header("FORCEREADY") + np.float64(energy).tobytes() + np.int32(natom).tobytes() + force.tobytes()
                               + vir.tobytes() + np.int32(len(extra)).tobytes() + extra)
                self.hasdata = False
            else:
                raise Exception("wrong message from server")


TMIPIManager = TMIPIManger
