"""Sensor layout for SpatialAttention (drop-in for the reference's
speech_decoding/utils/layout.py:6-43).

The reference reads real montages through `mne` / `mne_bids` and the dataset on
disk, and fails hard when either is missing.  So does this module -- unless the
caller EXPLICITLY opts into a stand-in layout (tests, benchmarks, synthetic data),
which honours the same post-conditions (float32 (C,2), per-axis min-max normalised,
scaled to [0.1, 0.9]; layout.py:37-43):

  args.sensor_layout    : (C,2) array/tensor of raw 2-D positions, used as given, or
  args.layout_seed      : seed of a synthetic uniform layout (its presence is the opt-in;
  args.synthetic_layout : ... as is this flag), with args.num_channels sensors

A missing `mne` install or an unreadable dataset is never papered over with fake geometry:
SpatialAttention's Fourier tables and SpatialDropout's distances would silently be wrong
(and SpatialDropout.loc is not in the state_dict, so not even a checkpoint would repair it).
"""
import numpy as np
import torch


def _normalise(loc):
    loc = np.asarray(loc, dtype=np.float64)
    span = loc.max(axis=0) - loc.min(axis=0)
    loc = (loc - loc.min(axis=0)) / span            # layout.py:38
    loc = loc * 0.8 + 0.1                           # layout.py:41 (margin of 0.1 per side)
    return torch.from_numpy(loc.astype(np.float32))


def synthetic_layout(num_channels, seed=0):
    rng = np.random.RandomState(1234 + int(seed))
    return _normalise(rng.rand(int(num_channels), 2))


def _get(args, name, default=None):
    try:
        return getattr(args, name)
    except Exception:
        try:
            return args[name]
        except Exception:
            return default


def ch_locations_2d(args):
    explicit = _get(args, "sensor_layout")
    if explicit is not None:
        if isinstance(explicit, torch.Tensor):
            explicit = explicit.detach().cpu().numpy()
        return _normalise(explicit)
    dataset = _get(args, "dataset")
    seed = _get(args, "layout_seed")
    if seed is not None or _get(args, "synthetic_layout"):          # explicit opt-in to a stand-in layout
        n = _get(args, "num_channels")
        if n is None:
            n = {"Brennan2018": 60, "Gwilliams2022": 208}.get(dataset)
        if n is None:
            raise ValueError("synthetic sensor layout requested but neither num_channels nor a known dataset given (%r)" % (dataset,))
        return synthetic_layout(n, seed or 0)
    try:
        import mne                                     # noqa: F401
    except ImportError as e:
        raise ImportError("speech_decoding.utils.layout: the sensor layout of %r needs `mne` (and `mne_bids` + the dataset "
                          "under args.root_dir for Gwilliams2022), exactly like the reference (layout.py:1-32).  For "
                          "synthetic data pass args.sensor_layout, or args.layout_seed / args.synthetic_layout with "
                          "args.num_channels." % (dataset,)) from e
    return _from_mne(dataset, _get(args, "root_dir"))


def _from_mne(dataset, root_dir):
    import mne
    if dataset == "Brennan2018":                       # layout.py:9-18: easycap-M10 minus broken channel 29
        montage = mne.channels.make_standard_montage("easycap-M10")
        info = mne.create_info(ch_names=montage.ch_names, sfreq=512.0, ch_types="eeg")
        info.set_montage(montage)
        pos = mne.channels.find_layout(info, ch_type="eeg").pos[:, :2]
        pos = np.delete(pos, 28, axis=0)
    elif dataset == "Gwilliams2022":                   # layout.py:20-32: KIT layout of subject 01
        import mne_bids
        path = mne_bids.BIDSPath(subject="01", session="0", task="0", datatype="meg",
                                 root="%s/data/Gwilliams2022/" % root_dir)
        raw = mne_bids.read_raw_bids(path)
        pos = mne.channels.find_layout(raw.info, ch_type="meg").pos[:, :2]
    else:
        raise ValueError("unknown dataset %r (layout.py:34)" % (dataset,))
    return _normalise(pos)
