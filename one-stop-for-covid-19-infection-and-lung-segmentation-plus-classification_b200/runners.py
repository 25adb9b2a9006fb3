"""The reference's six runner names (Scripts/app.py:36-57) over the B200 engine.

The reference runners take no arguments, download the Kaggle data inside Colab and only print / plot
(SURVEY.md D5).  These keep the names and the training recipe (graph, loss, optimizer, batch size, split,
callbacks, threshold sweeps) but accept the arrays and RETURN the artefacts; with no arrays they run on
the seeded synthetic stand-in (synthetic.py).  Reference defects listed in SURVEY 0.2 are not reproduced
(e.g. the cross-validation runners re-initialise the model for every fold).

  holdout_runner_unet_infection_segmentation   task1_preprocessing_plus_unet_with_comments.py:6-1508
  holdout_runner_unetplusplus_infection_segmentation   task1_unet_plus_plus.py
  three_fold_/four_fold_runner_unet_infection_segmentation   task1_crossval_{3,4}folds_unet.py
  runner_lung_segmentation   task3_lung_segmentation_unet.py
  runner_classification      task2_covid19_classifcation.py
"""
import numpy as np

from . import graphs as G
from . import losses as LS
from . import model as M
from . import synthetic as S


def _default_data(task, n, size, seed):
    return S.make_slices(n, size, seed=seed, task=task)


def _segmentation_holdout(graph_fn, task, cts=None, masks=None, new_dim=224, epochs=80, batch_size=32, lr=0.0005,
                          cosine=False, precision="float16", n_synthetic=64, seed=1234, thresholds=None, verbose=1,
                          checkpoint_path=None, comm=None):
    from sklearn.model_selection import train_test_split
    if cts is None:
        cts, masks = _default_data(task, n_synthetic, new_dim, seed)
    cts, masks = np.asarray(cts), np.asarray(masks)
    x_train, x_valid, y_train, y_valid = train_test_split(cts, masks, test_size=0.3, random_state=42)      # T1H:762
    model = M.Model(graph=graph_fn(new_dim, cts.shape[-1]), precision=precision, comm=comm)
    model.compile(optimizer=M.Adam(lr=lr), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])                  # T1H:1053
    callbacks = []
    if checkpoint_path:
        callbacks.append(M.ModelCheckpoint(checkpoint_path, monitor='val_dice_coeff', verbose=verbose, mode='max',
                                           save_best_only=True))                                           # T1H:1046
    if cosine:
        callbacks.append(M.CosineAnnealingScheduler(T_max=7, eta_max=0.0005, eta_min=0.0001, verbose=verbose))  # T1H:996
    results = model.fit(x_train, y_train, batch_size=batch_size, epochs=epochs, validation_data=(x_valid, y_valid),
                        callbacks=callbacks, verbose=verbose)                                              # T1H:1059
    if checkpoint_path:
        model.load_weights(checkpoint_path)                                                                # T1H:1073
    score = model.evaluate(x_valid, y_valid, batch_size=batch_size)                                        # T1H:1101
    thresholds = thresholds if thresholds is not None else [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    sweep = model.threshold_sweep(x_valid, y_valid, thresholds, batch_size=batch_size)                     # T1H:1196-1343
    best = int(np.argmax(sweep["f1"]))
    return dict(model=model, history=results.history, val_loss=score[0], val_dice_coeff=score[1], sweep=sweep,
                best_threshold=float(sweep["threshold"][best]), best_dice=float(sweep["f1"][best]),
                best_iou=float(sweep["iou"][best]), x_valid=x_valid, y_valid=y_valid)


def holdout_runner_unet_infection_segmentation(cts=None, infections=None, **kw):
    return _segmentation_holdout(G.unet, "infection", cts, infections, **kw)


def holdout_runner_unetplusplus_infection_segmentation(cts=None, infections=None, **kw):
    return _segmentation_holdout(G.unetpp, "infection", cts, infections, **kw)


def runner_lung_segmentation(cts=None, lungs=None, **kw):
    return _segmentation_holdout(G.unet, "lung", cts, lungs, **kw)


def _kfold_runner(n_splits, epochs_per_fold, cts=None, infections=None, new_dim=224, batch_size=32, lr=0.0005,
                  precision="float16", n_synthetic=64, seed=1234, thresholds=None, verbose=1):
    from sklearn.model_selection import KFold
    if cts is None:
        cts, infections = _default_data("infection", n_synthetic, new_dim, seed)
    cts, infections = np.asarray(cts), np.asarray(infections)
    thresholds = thresholds if thresholds is not None else list(np.arange(0.30, 0.80, 0.05))               # CV4:1221
    kf = KFold(n_splits=n_splits, shuffle=True, random_state=42)                                           # CV4:1047
    folds = []
    for fold, (tr, te) in enumerate(kf.split(cts)):
        model = M.Model(graph=G.unet(new_dim, cts.shape[-1]), precision=precision, seed=42 + fold)         # fresh per fold
        model.compile(optimizer=M.Adam(lr=lr), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
        h = model.fit(cts[tr], infections[tr], batch_size=batch_size, epochs=epochs_per_fold[fold],
                      validation_data=(cts[te], infections[te]), verbose=verbose)
        sweep = model.threshold_sweep(cts[te], infections[te], thresholds, batch_size=batch_size)
        folds.append(dict(history=h.history, sweep=sweep))
    return dict(folds=folds, **cv_report([f["sweep"] for f in folds]))                                     # CV4:1272-1364


def cv_report(sweeps):
    """The cross-validation tables of the reference (task1_crossval_4folds_unet.py:1272-1364): per metric a DataFrame
    with the thresholds as rows and the split numbers 1..k as columns, the best value and best threshold of every
    split, the maximum over everything and the mean over all thresholds and splits (the README's headline numbers).
    `sweeps` = one Model.threshold_sweep result per fold (all thresholds from ONE forward pass per fold instead of the
    reference's recompile + evaluate per threshold and metric)."""
    import pandas as pd
    the_range = np.asarray(sweeps[0]["threshold"], dtype=np.float64)
    cols = list(range(1, len(sweeps) + 1))
    tables, best, best_threshold, maximum, mean = {}, {}, {}, {}, {}
    for key, name in (("f1", "dice"), ("iou", "iou"), ("precision", "precision"), ("recall", "recall")):
        total = np.transpose(np.array([np.asarray(sw[key], dtype=np.float64) for sw in sweeps]))
        df = pd.DataFrame(data=total, index=the_range, columns=cols)
        tables[name] = df
        best[name] = np.array(df.max(axis=0))
        best_threshold[name] = [float(the_range[int(np.argmax(df[c].to_numpy()))]) for c in cols]
        maximum[name] = float(np.max(total))
        mean[name] = float(df.mean().mean())
    mean.update(f1=mean["dice"])          # (key used by earlier callers)
    return dict(tables=tables, best_per_split=best, best_threshold_per_split=best_threshold, maximum=maximum, mean=mean)


def four_fold_runner_unet_infection_segmentation(cts=None, infections=None, epochs=80, **kw):
    return _kfold_runner(4, [epochs] * 4, cts, infections, **kw)


def three_fold_runner_unet_infection_segmentation(cts=None, infections=None, epochs=(80, 20, 20), **kw):
    return _kfold_runner(3, list(epochs), cts, infections, **kw)


def runner_classification(cts=None, y_label=None, new_dim=224, epochs=25, batch_size=32, lr=0.0005, precision="float16",
                          n_synthetic=128, seed=1234, verbose=1):
    from sklearn.metrics import roc_auc_score
    from sklearn.model_selection import StratifiedShuffleSplit
    from sklearn.utils import class_weight
    if cts is None:
        cts, y_label = _default_data("class", n_synthetic, new_dim, seed)
    cts, y_label = np.asarray(cts), np.asarray(y_label).reshape(-1, 1)
    sss = StratifiedShuffleSplit(n_splits=1, test_size=0.3, random_state=42)                               # T2:647
    tr, te = next(sss.split(cts, y_label))
    x_train, x_valid, y_train, y_valid = cts[tr], cts[te], y_label[tr], y_label[te]
    weights = class_weight.compute_class_weight(class_weight='balanced', classes=np.unique(y_train.ravel()),
                                                y=y_train.ravel())                                         # T2:801-803
    model = M.Model(graph=G.classifier(new_dim, cts.shape[-1]), precision=precision)
    model.compile(loss='binary_crossentropy', optimizer=M.Adam(lr=lr), metrics=[LS.f1])                    # T2:828
    roc = M.RocCallback(training_data=(x_train, y_train), validation_data=(x_valid, y_valid))              # T2:706
    h = model.fit(x_train, y_train, batch_size=batch_size, epochs=epochs, validation_data=(x_valid, y_valid),
                  callbacks=[roc], class_weight=weights, verbose=verbose)                                  # T2:834
    probs = model.predict(x_valid)                                                                         # T2:919
    auc = float(roc_auc_score(y_valid, probs)) if len(np.unique(y_valid)) > 1 else float("nan")
    out = dict(model=model, history=h.history, auroc=auc, probs=probs, y_valid=y_valid, class_weight=weights)
    for thr in (0.5, 0.81):                                                                                # T2:926-989
        pr = (probs > thr).astype(np.float64)
        tp = float((pr * y_valid).sum())
        prec, rec = tp / max(pr.sum(), 1e-12), tp / max(y_valid.sum(), 1e-12)
        out["metrics@%.2f" % thr] = dict(accuracy=float((pr == y_valid).mean()), precision=prec, recall=rec,
                                         f1=2 * prec * rec / max(prec + rec, 1e-12))
    return out


def cluster_experiment(model, cts, x_valid, y_valid, layer_name="conv2d_9", n_components=1000, batch_size=32, threshold=0.547):
    """The reference's "easy / hard CT" clustering experiment (T1H:1386-1496): bottleneck activations of every slice
    (`Model(inputs=model.input, outputs=model.get_layer(layer_name).output).predict`, flattened channel-major like the
    reference's rollaxis + flatten), PCA, KMeans(2, random_state=0) fitted on `cts` and applied to the validation set,
    then the validation metrics of each cluster (loss, F-score and IoU at the reference's threshold 0.547).
    The network passes (activation tap, evaluation) run on the GPU; PCA / KMeans are the reference's own scikit-learn
    calls on the host (its feature matrix is a few hundred MB: not part of the training / inference hot path)."""
    from sklearn.cluster import KMeans
    from sklearn.decomposition import PCA

    def features(x):
        out = model.intermediate(x, layer_name)                       # (N, h, w, C) on the host
        return np.ascontiguousarray(np.transpose(out, (0, 3, 1, 2))).reshape(len(out), -1)      # rollaxis(img, 2).flatten()

    data = features(cts)
    pca = PCA(n_components=min(n_components, data.shape[0], data.shape[1]))
    new_data = pca.fit_transform(data)
    kmeans = KMeans(n_clusters=2, random_state=0, n_init=10).fit(new_data)
    valid_labels = kmeans.predict(pca.transform(features(x_valid)))
    model.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss,
                  metrics=[LS.FScore(threshold=threshold), LS.IOUScore(threshold=threshold)])          # T1H:1475
    out = dict(explained_variance=float(np.sum(pca.explained_variance_ratio_)), train_labels=kmeans.labels_,
               valid_labels=valid_labels, score_all=model.evaluate(x_valid, y_valid, batch_size=batch_size))
    for k in (0, 1):
        sel = np.where(valid_labels == k)[0]
        out["score_cluster_%d" % k] = model.evaluate(x_valid[sel], y_valid[sel], batch_size=batch_size) if len(sel) else None
        out["count_cluster_%d" % k] = int(len(sel))
    return out


def load_cases(cases, new_dim=224):
    """File-driven front end of the runners (the reference reads its 20 Kaggle volumes in a loop, T1H:390-393):
    `cases` = iterable of (ct_path, lung_mask_path, infection_mask_path or None) NIfTI files -> (cts, lungs-free
    infections) arrays ready for the runners above, with the reference's empty-mask filter left to the caller."""
    from . import nifti
    xs, ys = [], []
    for ct, lung, inf in cases:
        x, y, _ = nifti.preprocess_case(ct, lung, inf, new_dim=new_dim)
        xs.append(x)
        if y is not None:
            ys.append(y)
    return np.concatenate(xs), (np.concatenate(ys) if ys else None)


RUNNERS = {
    "one": three_fold_runner_unet_infection_segmentation, "two": four_fold_runner_unet_infection_segmentation,
    "three": holdout_runner_unet_infection_segmentation, "four": holdout_runner_unetplusplus_infection_segmentation,
    "five": runner_classification, "six": runner_lung_segmentation,
}
