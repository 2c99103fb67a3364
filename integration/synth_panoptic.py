"""Synthetic ``panoptic_*`` dataset classes for the "tools run unchanged" harness: no dataset is in
the container, so the unmodified ``tools/evaluate.py`` / ``tools/train_3d.py`` are pointed (through the YAML's
``DATASET.TEST_DATASET`` / ``TRAIN_DATASET``) at these classes, which honour the reference's item contracts --
the 6-tuple of ``lib/dataset/panoptic.py:269-288`` and the 18-tuple (three view sets) of
``lib/dataset/JointsDatasetSSV.py:615-640`` -- and the ``evaluate`` return shape of ``lib/dataset/panoptic.py:385-390``.
Their names contain "panoptic" (the branches at ``lib/core/function.py:72,246,369,440``).  HARNESS ONLY: part of the
demo / tests, not of the backend."""
import numpy as np
import torch
from torch.utils.data import Dataset

from selfpose3d_b200 import synthetic


def _unbatch(x):
    if isinstance(x, torch.Tensor):
        return x[0]
    if isinstance(x, dict):
        return {k: _unbatch(v) for k, v in x.items()}
    return x


class SyntheticPanoptic(Dataset):
    """``dataset.panoptic_synth(cfg, image_set, is_train, transform)``: a few frames of a 5-camera synthetic scene."""
    LENGTH = 2

    def __init__(self, cfg, image_set, is_train, transform=None):
        self.cfg, self.is_train = cfg, is_train
        self.image_size = [int(v) for v in cfg.NETWORK.IMAGE_SIZE]
        self.heatmap_size = [int(v) for v in cfg.NETWORK.HEATMAP_SIZE]
        self.num_joints = int(cfg.NETWORK.NUM_JOINTS)
        self.num_views = int(cfg.DATASET.CAMERA_NUM)
        self.max_people = int(cfg.MULTI_PERSON.MAX_PEOPLE_NUM)
        self.cube = [int(v) for v in cfg.MULTI_PERSON.INITIAL_CUBE_SIZE]
        self.cams = synthetic.ring_cameras(self.num_views, seed=0)

    def __len__(self):
        return self.LENGTH

    def _people(self, idx):
        return synthetic.synthetic_people(1, seed=100 + idx, num_joints=self.num_joints)

    def _frame(self, idx, rotation=0.0, scale_mul=1.0, image_seed=0):
        V, J = self.num_views, self.num_joints
        meta = synthetic.make_meta(self.cams, 1, self.image_size, rotation=[[rotation]] * V, scale_mul=[[scale_mul]] * V)
        people = self._people(idx)
        hms = synthetic.render_heatmaps(people, meta, self.image_size, self.heatmap_size, num_joints=J, sigma=2.0)
        images = synthetic.random_images(1, V, self.image_size, seed=idx * 7 + image_seed)
        joints = np.zeros((self.max_people, J, 3))
        n = min(len(people[0]), self.max_people)
        joints[:n] = np.asarray(people[0])[:n]
        rs = np.random.RandomState(idx)
        for v, m in enumerate(meta):
            m["num_person"] = torch.tensor([n])
            m["joints_3d"] = torch.from_numpy(joints)[None]
            m["joints_3d_vis"] = torch.ones(1, self.max_people, J, 3, dtype=torch.float64)
            m["roots_3d"] = torch.from_numpy(joints[:, int(self.cfg.DATASET.ROOTIDX)])[None]
            m["joints"] = torch.from_numpy(rs.uniform(5, min(self.image_size) - 5, (1, self.max_people, J, 2)))
            m["joints_vis"] = torch.ones(1, self.max_people, J, 2, dtype=torch.float64)
            m["mis_count"] = torch.tensor([0])
            m["image"] = "synthetic_%03d_view%d" % (idx, v)
        target_3d = torch.rand(1, *self.cube, generator=torch.Generator().manual_seed(idx))
        inputs = [_unbatch(x) for x in images]
        targets_2d = [_unbatch(h) for h in hms]
        weights_2d = [torch.ones(J, 1) for _ in range(V)]
        targets_3d = [target_3d[0] for _ in range(V)]
        metas = [_unbatch(m) for m in meta]
        return inputs, targets_2d, weights_2d, targets_3d, metas, [h.clone() for h in targets_2d]

    def __getitem__(self, idx):
        return self._frame(idx)

    def evaluate(self, preds, roots, output_dir=None):
        """((aps, aps_root), (recs, recs_root), (mpjpe, mpjpe_root), (recall, recall_root)) -- MPJPE-thresholded
        precision of the valid predictions against the synthetic ground truth (lib/dataset/panoptic.py:290-390)."""
        root_idx = int(self.cfg.DATASET.ROOTIDX)
        thresholds = np.arange(25, 155, 25)
        errs, errs_root, total_gt = [], [], 0
        for i, (pred, root) in enumerate(zip(preds, roots)):
            gt = np.asarray(self._people(i % self.LENGTH)[0])
            total_gt += len(gt)
            for p, r in zip(pred, root):
                if p[0, 3] < 0:
                    continue
                errs.append(min(float(np.sqrt(((p[:, :3] - g) ** 2).sum(-1)).mean()) for g in gt))
                errs_root.append(min(float(np.sqrt(((r[:3] - g[root_idx]) ** 2).sum())) for g in gt))

        def table(e):
            e = np.asarray(e) if e else np.zeros(0)
            aps = [float((e < t).sum()) / max(len(e), 1) for t in thresholds]
            recs = [float((e < t).sum()) / max(total_gt, 1) for t in thresholds]
            ok = e[e < 500.0]
            mean = float(ok.mean()) if len(ok) else (float(e.mean()) if len(e) else 0.0)
            return aps, recs, mean, float((e < 500.0).sum()) / max(total_gt, 1)

        a, r, m, c = table(errs)
        ar, rr, mr, cr = table(errs_root)
        return (a, ar), (r, rr), (m, mr), (c, cr)


class SyntheticPanopticSSV(SyntheticPanoptic):
    """``dataset.panoptic_synth_ssv``: training items are three view sets (two augmented, one plain) with the ``meta``
    entries the self-supervised forward reads (``trans``, ``hflip``, ``camera.f/c``, pseudo 2-D ``joints``)."""
    AUGMENT = ((12.0, 1.1, False), (-8.0, 0.9, True), (0.0, 1.0, False))

    def __getitem__(self, idx):
        if not self.is_train:
            return self._frame(idx)
        from selfpose3d_b200.utils.transforms import get_affine_transform
        out = []
        for s, (rot, mul, flip) in enumerate(self.AUGMENT):
            item = self._frame(idx, rotation=rot, scale_mul=mul, image_seed=s + 1)
            metas = item[4]
            for m in metas:
                cam = {k: v.float() for k, v in m["camera"].items()}
                cam["f"] = torch.stack([cam["fx"], cam["fy"]]).reshape(2, 1)
                cam["c"] = torch.stack([cam["cx"], cam["cy"]]).reshape(2, 1)
                m["camera"] = cam
            trans = get_affine_transform(metas[0]["center"].numpy(), metas[0]["scale"].numpy(), float(metas[0]["rotation"]),
                                         self.image_size)
            metas[0]["trans"] = torch.from_numpy(trans.astype(np.float32))
            metas[0]["hflip"] = torch.tensor(flip)
            out += list(item)
        return tuple(out)


def register(dataset_module):
    """Make the classes resolvable as ``dataset.panoptic_synth`` / ``dataset.panoptic_synth_ssv`` (tools/train_3d.py:93,113
    evaluate ``'dataset.' + cfg.DATASET.TEST_DATASET``)."""
    dataset_module.panoptic_synth = SyntheticPanoptic
    dataset_module.panoptic_synth_ssv = SyntheticPanopticSSV
