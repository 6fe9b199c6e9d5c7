"""Physical constants and element tables under the names the reference exports
(TensorMol/PhysicalData.py).  Units: Angstrom, Hartree, fs; MD works in SI per mole."""
from __future__ import annotations

import numpy as np
from math import pi as Pi

atoi = {'H': 1, 'He': 2, 'Li': 3, 'Be': 4, 'B': 5, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'Ne': 10, 'Na': 11, 'Mg': 12, 'Al': 13, 'Si': 14, 'P': 15,
        'S': 16, 'Cl': 17, 'Ar': 18, 'K': 19, 'Ca': 20, 'Br': 35, 'Cs': 55, 'Pb': 82}
itoa = {v: k for k, v in atoi.items()}

# Grimme D2 parameters used by the vdW term (reference PhysicalData.py:25-26): radius in Angstrom, C6 in J nm^6 / mol
atomic_vdw_radius = {1: 1.001, 2: 1.012, 3: 0.825, 4: 1.408, 5: 1.485, 6: 1.452, 7: 1.397, 8: 1.342, 9: 1.287, 10: 1.243}
C6_coff = {1: 0.14, 2: 0.08, 3: 1.16, 4: 1.61, 5: 3.13, 6: 1.75, 7: 1.23, 8: 0.70, 9: 0.75, 10: 0.63}

ATOMICMASSESAMU = np.array([1.00794, 4.002602, 6.941, 9.012182, 10.811, 12.0107, 14.0067, 15.9994, 18.9984032, 20.1791, 22.98976928,
                            24.3050, 26.9815386, 28.0855, 30.973762, 32.065, 35.453, 39.948, 39.0983, 40.078, 44.955912, 47.867, 50.9415,
                            51.9961, 54.938045, 55.845, 58.933195, 58.6934, 63.546, 65.38, 69.723, 72.63, 74.92160, 78.96, 79.904, 83.798])
ATOMICMASSES = 0.000999977 * ATOMICMASSESAMU          # kg/mol

GOLDENRATIO = (np.sqrt(5.) + 1.0) / 2.0
KAYBEETEE = 0.000950048
BOHRPERA = 1.889725989
ANGSTROMPERMETER = pow(10.0, 10.0)
BOHRPERM = BOHRPERA * ANGSTROMPERMETER
BOHRINM = 0.52917720859 * pow(10.0, -10.0)
KJPERHARTREE = 2625.499638
JOULEPERHARTREE = KJPERHARTREE * 1000.0
JOULEPERKCAL = 4183.9953
KCALPERHARTREE = 627.509474
WAVENUMBERPERHARTREE = 219474.63
ELECTRONPERPROTONMASS = 1836.15267
FEMTOPERUNIT = pow(10.0, -15.0)
PICOPERUNIT = pow(10.0, -12.0)
SPEEDOFLIGHT = 299792458.0
FSPERAU = 0.0241888
AVOCONST = 6.02214086 * np.power(10.0, 23.0)
AUPERDEBYE = 0.393456
IDEALGASR = 8.3144621
AMUINKG = 1.660538782 * pow(10.0, -27.0)
SECPERATOMIC = 2.418884326505 * pow(10.0, -17.0)
KCONVERT = (4.359744 * pow(10.0, -18.0)) / (BOHRINM * BOHRINM * AMUINKG)
CMCONVERT = 1.0 / (2.0 * Pi * SPEEDOFLIGHT * 100.0)


def AtomicNumber(symbol):
    return atoi[symbol]


def AtomicSymbol(number):
    return itoa[int(number)]
