"""The data format on the CALLER side of the hot path: how the reference's loader turns two
streams of (strong view, weak view) pairs into the 4-tuple `PTrainer.run_step` consumes,
`(label_strong, label_weak, unlabel_strong, unlabel_weak)` = lists of dicts with "image" (uint8 CHW), "height",
"width" and, for labelled data, "instances" (`pt/data/common.py:106-180`, built by
`pt/data/build.py:build_semisup_batch_data_loader_two_crop`). Augmentation itself (`pt/data/dataset_mapper.py`,
PIL / torchvision transforms in DataLoader worker processes) stays outside the scope; any iterable yielding
`(strong_dict, weak_dict)` pairs can be plugged in.

The grouping keeps the reference's exact emission order, including its quirk: while ONE of the two buckets in use
is already full, further items of that stream are DROPPED (the two streams are advanced in lock step by `zip`,
`pt/data/common.py:145-163`)."""


class AspectRatioGroupedSemiSupDatasetTwoCrop:
    """`pt/data/common.py:106-180`. dataset = (labelled iterable, unlabelled iterable), each yielding
    `(strong, weak)` pairs of dicts with "width" / "height"; batch_size = (labelled, unlabelled) per iteration.
    Images with w > h and w <= h are batched separately (less padding), independently for the two streams."""

    def __init__(self, dataset, batch_size):
        self.label_dataset, self.unlabel_dataset = dataset
        self.batch_size_label, self.batch_size_unlabel = batch_size[0], batch_size[1]
        # [orientation][0 = strong (q), 1 = weak (k)]
        self._label_buckets = [([], []) for _ in range(2)]
        self._unlabel_buckets = [([], []) for _ in range(2)]

    @staticmethod
    def _bucket_id(d):
        return 0 if d["width"] > d["height"] else 1

    def __iter__(self):
        lab = unl = None  # the (strong, weak) bucket each stream filled last
        for d_label, d_unlabel in zip(self.label_dataset, self.unlabel_dataset):
            if lab is None or len(lab[0]) != self.batch_size_label:
                lab = self._label_buckets[self._bucket_id(d_label[0])]
                lab[0].append(d_label[0])
                lab[1].append(d_label[1])
            if unl is None or len(unl[0]) != self.batch_size_unlabel:
                unl = self._unlabel_buckets[self._bucket_id(d_unlabel[0])]
                unl[0].append(d_unlabel[0])
                unl[1].append(d_unlabel[1])
            if len(lab[0]) == self.batch_size_label and len(unl[0]) == self.batch_size_unlabel:
                yield lab[0][:], lab[1][:], unl[0][:], unl[1][:]
                for b in (lab[0], lab[1], unl[0], unl[1]):
                    del b[:]


def build_semisup_batch_loader_two_crop(label_pairs, unlabel_pairs, batch_label, batch_unlabel, world_size=1):
    """`pt/data/build.py` build_semisup_batch_data_loader_two_crop: total batch sizes are divided by the world
    size (each rank draws its own shard of the two streams) and grouped by aspect ratio. Returns an iterator of
    4-tuples for `PTrainer(cfg, data_loader_iter=...)`."""
    if batch_label <= 0 or batch_label % world_size or batch_unlabel <= 0 or batch_unlabel % world_size:
        raise AssertionError(f"Total batch sizes ({batch_label}, {batch_unlabel}) must be positive and divisible "
                             f"by the number of gpus ({world_size}).")
    return iter(AspectRatioGroupedSemiSupDatasetTwoCrop(
        (label_pairs, unlabel_pairs), (batch_label // world_size, batch_unlabel // world_size)))
