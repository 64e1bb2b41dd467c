"""Host-side preparation of the labelled pixel lists and shadow statistics that the data-set loaders hand to the
importers (reference: common/common_nn_ops.py:455-543, names kept; re-exported from hypelcnn_b200.common.common_nn_ops).
Vectorised numpy; the stratified splits use scikit-learn exactly like the reference so that seeded splits coincide.
Pinned by tests/golden/sample_ops_golden.npz, produced by running the reference's own functions."""
import numpy

INVALID_TARGET_VALUE = 255


def read_targets_from_image(targets, class_range):
    """Label raster [H, W] -> rows (x, y, class), classes in class_range order, pixels of a class in row-major order."""
    targets = numpy.asarray(targets)
    blocks = [numpy.empty((0, 3), dtype=int)]
    for cls in class_range:
        ys, xs = numpy.nonzero(targets == cls)
        blocks.append(numpy.stack([xs.astype(int), ys.astype(int), numpy.full(xs.shape, cls, dtype=int)], axis=1))
    return numpy.concatenate(blocks, axis=0)


def create_target_image_via_samples(sample_set, scene_shape):
    """Rasterise the three target lists; a pixel listed twice keeps the class of its LAST occurrence in
    (training, test, validation) order, pixels never listed are INVALID_TARGET_VALUE."""
    rows = numpy.concatenate([numpy.asarray(t).reshape(-1, 3) for t in
                              (sample_set.training_targets, sample_set.test_targets, sample_set.validation_targets)]).astype(int)
    image = numpy.full((scene_shape[0], scene_shape[1]), INVALID_TARGET_VALUE, dtype=numpy.uint8)
    if rows.shape[0]:
        flat = rows[:, 1] * scene_shape[1] + rows[:, 0]
        _, last = numpy.unique(flat[::-1], return_index=True)           # first hit in the reversed list = last occurrence
        keep = rows.shape[0] - 1 - last
        image.reshape(-1)[flat[keep]] = rows[keep, 2]
    return image


def create_colored_image(target_image, color_list):
    """Class raster -> RGB raster through the loader's colour table; classes beyond the table stay black."""
    target_image = numpy.asarray(target_image)
    colors = numpy.asarray(color_list, dtype=numpy.uint8).reshape(-1, 3)
    out = numpy.zeros(target_image.shape[:2] + (3,), dtype=numpy.uint8)
    known = target_image < len(colors)
    out[known] = colors[target_image[known].astype(numpy.intp)]
    return out


def calculate_shadow_ratio(casi, shadow_map, shadow_map_inverse):
    """Per band: mean over the pixels marked in shadow_map_inverse (lit) / mean over the pixels marked in shadow_map,
    float32.  The maps mark pixels with non-zero entries."""
    casi = numpy.asarray(casi)
    in_shadow, lit = numpy.asarray(shadow_map) != 0, numpy.asarray(shadow_map_inverse) != 0
    mean_lit = casi[lit].mean(axis=0, dtype=numpy.float64)
    mean_shadow = casi[in_shadow].mean(axis=0, dtype=numpy.float64)
    return (mean_lit / mean_shadow).astype(numpy.float32)


def _stratified(rows, **split_args):
    from sklearn.model_selection import StratifiedShuffleSplit
    first, second = next(StratifiedShuffleSplit(n_splits=1, **split_args).split(rows[:, 0:1], rows[:, 2]))
    return rows[first], rows[second]


def shuffle_training_data_using_ratio(result, train_data_ratio):
    """-> (train_set, validation_set): class-stratified, train_data_ratio of the rows for training (unseeded, like the
    reference: every run draws a new split)."""
    return _stratified(result, train_size=train_data_ratio)


def shuffle_test_data_using_ratio(train_set, test_data_ratio):
    """-> (test_set, train_set): class-stratified with random_state 0 (reproducible); ratio 0 keeps everything."""
    if not test_data_ratio > 0:
        return numpy.empty([0, train_set.shape[1]]), train_set
    remaining, test = _stratified(train_set, test_size=test_data_ratio, random_state=0)
    return test, remaining


def shuffle_training_data_using_size(class_count, result, train_data_size, validation_size):
    """Per class: train_data_size random rows for training (9/10 of the class when it is smaller than that), the
    rest — or validation_size random rows of the rest — for validation.  Draws from numpy's global generator in
    the reference's order, so numpy.random.seed makes both agree."""
    labels = result[:, 2]
    train_parts = [numpy.empty([0, result.shape[1]], dtype=int)]
    validation_parts = [numpy.empty([0, result.shape[1]], dtype=int)]
    for cls in class_count:
        members = numpy.flatnonzero(labels == cls)
        n = members.shape[0]
        if n == 0:
            continue
        take = (n * 9) // 10 if n < train_data_size else train_data_size
        chosen = numpy.random.choice(n, take, replace=False)
        rest = numpy.setdiff1d(numpy.arange(n), chosen)                  # ascending, as the reference's filter leaves it
        if validation_size is not None:
            validation_size = min(validation_size, rest.shape[0])        # the shrunken size carries over to later classes
            rest = rest[numpy.random.choice(rest.shape[0], validation_size, replace=False)]
        train_parts.append(result[members[chosen], :])
        validation_parts.append(result[members[rest], :])
    return numpy.vstack(train_parts), numpy.vstack(validation_parts)
