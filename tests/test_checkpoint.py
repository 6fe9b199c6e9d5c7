"""CPU tests of the TensorFlow-free checkpoint reader (tensormol_b200/TFNetworks/TFCheckpoint.py).

No TensorFlow-written file exists offline, so the reader is exercised against the module's own writer plus hand-checked
pieces of the published formats: CRC-32C known answers (RFC 3720 B.4), the LevelDB table footer / trailer layout, protobuf
wire encodings, and Python-2-style pickles of stand-in classes living under the reference's module paths."""
import os
import pickle
import struct
import sys
import types

import numpy as np
import pytest

from tensormol_b200.TFNetworks import TFCheckpoint as ck


def test_crc32c_known_answers():
    assert ck.crc32c(b"123456789") == 0xE3069283                    # the standard check value of CRC-32C
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43               # ... of ones
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E                 # ... incrementing
    c = ck.crc32c(b"foo")
    assert ck.masked_crc32c(b"foo") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def test_varint_and_entry_encoding():
    assert ck._put_varint(300) == b"\xac\x02" and ck._get_varint(b"\xac\x02", 0) == (300, 2)
    raw = ck._encode_entry(1, (768, 200), 0, 4096, 768 * 200 * 4, 0xDEADBEEF)
    # dtype: field 1 varint; shape: field 2 message of dim messages; offset 4; size 5; crc32c field 6 fixed32
    assert raw[:2] == b"\x08\x01" and raw[2] == 0x12
    assert raw.endswith(b"\x35" + struct.pack("<I", 0xDEADBEEF))
    e = ck._parse_entry(raw)
    assert e == dict(dtype=1, shape=(768, 200), shard_id=0, offset=4096, size=768 * 200 * 4, crc32c=0xDEADBEEF)


def _random_weights(eles, inshape, hidden, seed):
    rng = np.random.default_rng(seed)
    out = {"charge": {}, "energy": {}}
    for net in out:
        for z in eles:
            fan, layers = inshape, []
            for h in list(hidden) + [1]:
                layers.append((rng.standard_normal((fan, h)), rng.standard_normal(h)))
                fan = h
            out[net][z] = layers
    return out


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bundle_round_trip_with_reference_variable_names(tmp_path, dtype):
    eles, hidden, inshape = [1, 6, 7, 8], [40, 24, 16], 768
    W = _random_weights(eles, inshape, hidden, 0)
    variables = ck.variables_from_weights(W, dtype)
    assert "EnergyNet/8_hidden1/weights" in variables and "EnergyNet/8_hidden2/biaseslayer1" in variables
    assert "DipoleNet/1_hidden3_charge/biases" in variables and "DipoleNet/6_regression_linear_charge/weights" in variables
    # what a Saver also stores: optimiser slots and counters under other names -- must be ignored
    variables["EnergyNet/8_hidden1/weights/Adam"] = np.zeros((inshape, 40), dtype)
    variables["EnergyNet/8_hidden1/weights/Adam_1"] = np.zeros((inshape, 40), dtype)
    variables["beta1_power"] = np.array(0.9, np.float32)
    variables["global_step"] = np.array(12345, np.int64)
    prefix = str(tmp_path / "net" / "Mol_x-chk-500")
    ck.write_bundle(prefix, variables)
    # file layout: footer = 40 bytes of handles + the table magic, little endian
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xdb4775248b80fb57
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(np.asarray(v).nbytes for v in variables.values())
    header, entries = ck.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and len(entries) == len(variables)          # several 4 KiB blocks: 68 entries
    assert entries["global_step"]["shape"] == () and entries["global_step"]["dtype"] == 9
    back = ck.read_bundle(prefix, verify=True)
    assert set(back) == set(variables)
    for k, v in variables.items():
        assert back[k].dtype == np.asarray(v).dtype and np.array_equal(back[k], v)
    W2 = ck.weights_from_variables(back, eles, hidden, inshape)
    for net in W:
        for z in eles:
            for (a, b), (c, d) in zip(W[net][z], W2[net][z]):
                assert c.dtype == np.float64 and np.array_equal(a.astype(dtype), c) and np.array_equal(b.astype(dtype), d)
    assert ck.latest_checkpoint(str(tmp_path / "net")) == prefix
    # an outer scope in front of every name is resolved; a wrong shape is refused
    scoped = {"tower_0/" + k: v for k, v in back.items()}
    assert np.array_equal(ck.weights_from_variables(scoped, eles, hidden, inshape)["energy"][7][0][0], W2["energy"][7][0][0])
    with pytest.raises(ck.CheckpointError):
        ck.weights_from_variables(back, eles, [40, 24, 17], inshape)
    with pytest.raises(ck.CheckpointError):
        ck.weights_from_variables(back, [1, 8, 17], hidden, inshape)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c-chk-1")
    ck.write_bundle(prefix, {"a/weights": np.arange(12.0).reshape(3, 4), "a/biases": np.ones(4)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.read_bundle(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ck.CheckpointError):
        ck.read_index(prefix + ".index")
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_index(prefix + ".index")


def _fake_reference_modules():
    """Classes under the reference's module paths, as its pickles name them (instances of classes this package lacks)."""
    mods = {}
    for name in ("RefTM", "RefTM.TFNetworks", "RefTM.TFNetworks.TFMolInstanceDirect", "RefTM.Containers", "RefTM.Containers.TensorMolData"):
        mods[name] = types.ModuleType(name)
    inst_cls = type("MolInstance_DirectBP_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout", (object,), {"__module__": "RefTM.TFNetworks.TFMolInstanceDirect"})
    data_cls = type("TensorMolData_BP_Direct_EE_WithEle", (object,), {"__module__": "RefTM.Containers.TensorMolData"})
    setattr(mods["RefTM.TFNetworks.TFMolInstanceDirect"], inst_cls.__name__, inst_cls)
    setattr(mods["RefTM.Containers.TensorMolData"], data_cls.__name__, data_cls)
    return mods, inst_cls, data_cls


def test_manager_and_instance_pickles_lead_to_the_checkpoint(tmp_path):
    nets = str(tmp_path) + "/"
    inst_name = "Mol_set_ANI1_Sym_Direct_fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout_act_sigmoid100"
    os.makedirs(nets + inst_name)
    prefix = nets + inst_name + "/" + inst_name + "-chk-420"
    W = _random_weights([1, 8], 256, [8, 8, 8], 1)
    ck.write_bundle(prefix, ck.variables_from_weights(W))
    ck.write_bundle(nets + inst_name + "/" + inst_name + "-chk-100", {"x": np.zeros(1)})
    mods, inst_cls, data_cls = _fake_reference_modules()
    sys.modules.update(mods)
    try:
        tdata = data_cls()
        tdata.eles = [1, 8]
        tdata.name = "set"
        tdata.scratch = np.arange(4)
        inst = inst_cls()
        inst.__dict__.update(dict(name=inst_name, HiddenLayers=[8, 8, 8], eles=[1, 8], inshape=256, TData=tdata,
                                  chk_file="./networks/" + inst_name + "/" + inst_name + "-chk-420"))
        with open(nets + inst_name + ".tfn", "wb") as fh:
            pickle.dump(inst.__dict__, fh, protocol=2)
        with open(nets + "water_network.tfm", "wb") as fh:
            pickle.dump(dict(name="water_network", NetType="fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout",
                             TrainedNetworks=[inst_name], TData=tdata, Instances=inst, n_train=500), fh, protocol=2)
    finally:
        for m in mods:
            sys.modules.pop(m, None)
    # the classes are gone now: the loader must not need them
    mgr = ck.load_tm_pickle(nets + "water_network.tfm")
    assert mgr["TrainedNetworks"] == [inst_name] and mgr["TData"].eles == [1, 8] and np.array_equal(mgr["TData"].scratch, np.arange(4))
    assert type(mgr["Instances"]).__name__.startswith("MolInstance_DirectBP_EE") and mgr["Instances"].HiddenLayers == [8, 8, 8]
    chk, state = ck.find_reference_network("water_network", nets)
    assert chk == prefix and state["HiddenLayers"] == [8, 8, 8]
    w = ck.weights_from_variables(ck.read_bundle(chk), state["eles"], state["HiddenLayers"], state["inshape"])
    assert np.array_equal(w["charge"][8][3][0], W["charge"][8][3][0])
    # no .tfn: the newest checkpoint of the instance directory
    os.remove(nets + inst_name + ".tfn")
    chk2, _ = ck.find_reference_network("water_network", nets)
    assert chk2 == prefix
    assert ck.find_reference_network("absent", nets) == (None, None)


def test_manager_follows_the_reference_load_chain_without_a_gpu(tmp_path, monkeypatch):
    """TFMolManage(Name_) -> <Name_>.tfm -> <instance>.tfn -> chk_file -> weights handed to the engine (the CUDA engine is
    replaced by a recorder here; the GPU version of this test evaluates with the restored weights)."""
    from tensormol_b200 import PARAMS, Mol, MolDigester, MSet, TensorMolData_BP_Direct_EE_WithEle
    import importlib
    tmm = importlib.import_module("tensormol_b200.TFNetworks.TFMolManage")   # the module, not the re-exported class

    class FakeEngine:
        elu_shift = elu_alpha = 0.0

        def __init__(self, eles, hidden, params, device=0):
            self.eles, self.hidden, self.set = list(eles), list(hidden), None

        def set_gemm_mode(self, mode):
            pass

        def set_weights(self, w):
            self.set = w

    monkeypatch.setattr(tmm, "Engine", FakeEngine)
    net = "fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout"
    nets = str(tmp_path) + "/"
    inst = "Mol_t_ANI1_Sym_Direct_" + net
    W = _random_weights([1, 8], 256, [12, 10, 8], 3)
    ck.write_bundle(nets + inst + "/" + inst + "-chk-7", ck.variables_from_weights(W, np.float32))
    with open(nets + inst + ".tfn", "wb") as fh:
        pickle.dump(dict(name=inst, HiddenLayers=[12, 10, 8], eles=[1, 8], chk_file="./networks/" + inst + "/" + inst + "-chk-7"), fh, protocol=2)
    with open(nets + "water_network.tfm", "wb") as fh:
        pickle.dump(dict(name="water_network", NetType=net, TrainedNetworks=[inst]), fh, protocol=2)
    monkeypatch.setitem(PARAMS, "networks_directory", nets)
    monkeypatch.setitem(PARAMS, "HiddenLayers", [12, 10, 8])
    monkeypatch.setitem(PARAMS, "EECutoffOn", 0)
    a = MSet("t", center_=False)
    a.mols = [Mol(np.array([1, 1, 8], np.uint8), np.array([[0.757, 0.586, 0.0], [-0.757, 0.586, 0.0], [0.0, 0.0, 0.0]]))]
    tset = TensorMolData_BP_Direct_EE_WithEle(a, MolDigester(a.AtomTypes(), name_="ANI1_Sym_Direct", OType_="EnergyAndDipole"), order_=1,
                                              num_indis_=1, type_="mol", WithGrad_=True)
    manager = tmm.TFMolManage("water_network", tset, False, net, False, False)
    got = manager.Instances.engine.set
    assert got is not None and manager.TrainedNetworks == [inst]
    for n in W:
        for z in W[n]:
            for (w0, b0), (w1, b1) in zip(W[n][z], got[n][z]):
                assert np.array_equal(w0.astype(np.float32), w1) and np.array_equal(b0.astype(np.float32), b1)
    # save under the reference's names and read back
    manager.SaveCheckpoint(nets + "again-chk-1")
    assert set(ck.read_bundle(nets + "again-chk-1")) == set(ck.variables_from_weights(W))
    monkeypatch.setitem(PARAMS, "HiddenLayers", [12, 10, 9])
    with pytest.raises(ck.CheckpointError, match="HiddenLayers"):
        tmm.TFMolManage("water_network", tset, False, net, False, False)
