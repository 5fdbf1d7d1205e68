"""trainer.py -- drop-in for the reference's hyper-parameter inference script (reference trainer.py:17-92).

Same two entry points, same CSV layout: `train_sfgp(name)` reads `Data/{name}_hifi_train.csv`, `train_mfgp(name)` reads
`Data/{name}_lofi_train.csv` + `Data/{name}_hifi_train.csv` (columns X, Y, f; one header row), both fit the model by
minimising the negative log marginal likelihood with L-BFGS-B from the reference's initial length scales (0.01; 0.5 / 0.1)
and write the LOG-scaled hyper-parameters to `Data/{name}_sf_hyp.csv` / `Data/{name}_mf_hyp.csv` with the reference's column
labels -- the files simulator.py consumes.  The likelihood and its gradient are evaluated on the device
(gaussian_process.likelihood_and_grad -> mfgp_nlml_grad); the reference differentiates with autograd on the CPU.

The reference asks "Save ... hyperparameters?" on stdin before writing (:47, :87); `save=None` keeps that prompt,
`save=True/False` answers it programmatically.
"""
import numpy as np
import pandas as pd

from .gaussian_process import MFGP, SFGP

SF_LABELS = ['mu_sf', 's^2_sf', 'L_sf', 'noise_sf']
MF_LABELS = ['mu_lo', 's^2_lo', 'L_lo', 'mu_hi', 's^2_hi', 'L_hi', 'rho', 'noise_lo', 'noise_hi']


def _report_and_save(model, labels, question, path, save):
    hyp = model.hyp
    ehyp = np.exp(model.hyp)
    for i in range(len(labels)):
        print(f"{labels[i]} = {ehyp[i]} // log({labels[i]}) = {hyp[i]}")
    valid = input(question) if save is None else ("y" if save else "n")
    if valid.lower() == "y":
        hyp_df = pd.DataFrame(model.hyp.reshape(1, -1))
        hyp_df.columns = labels
        hyp_df.to_csv(path, index=False)


def train_sfgp(name, data_dir="Data", save=None, callback=True):
    """reference trainer.py:17-52."""
    sifi = np.loadtxt(f"{data_dir}/{name}_hifi_train.csv", skiprows=1, delimiter=',')     # train from hifi only (:27)
    X = sifi[:, [0, 1]].reshape(-1, 2)
    y = sifi[:, [2]].reshape(-1, 1)
    len_sf = 0.01
    model = SFGP(X, y, len_sf)
    model.train(callback=callback)
    _report_and_save(model, SF_LABELS, "Save single-fidelity hyperparameters?", f"{data_dir}/{name}_sf_hyp.csv", save)
    return model


def train_mfgp(name, data_dir="Data", save=None, callback=True):
    """reference trainer.py:55-92."""
    lofi = np.loadtxt(f"{data_dir}/{name}_lofi_train.csv", skiprows=1, delimiter=',')
    hifi = np.loadtxt(f"{data_dir}/{name}_hifi_train.csv", skiprows=1, delimiter=',')
    X_L = lofi[:, [0, 1]].reshape(-1, 2)
    y_L = lofi[:, 2].reshape(-1, 1)
    X_H = hifi[:, [0, 1]].reshape(-1, 2)
    y_H = hifi[:, 2].reshape(-1, 1)
    len_L = 0.5
    len_H = 0.1
    model = MFGP(X_L, y_L, X_H, y_H, len_L, len_H)
    model.train(callback=callback)
    _report_and_save(model, MF_LABELS, "Save multi-fidelity hyperparameters?", f"{data_dir}/{name}_mf_hyp.csv", save)
    return model


if __name__ == "__main__":
    np.random.seed(1234)        # reference trainer.py:99-103
    name = "australia9"
    train_sfgp(name)
    train_mfgp(name)
