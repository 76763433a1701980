from scipy.io import wavfile


def read(path, dtype="int16", **_):
    rate, data = wavfile.read(path)
    return data.astype(dtype), rate
