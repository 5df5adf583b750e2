"""hparams surface of the reference (hparams.py:1-114): same names and values for everything the forward path and
the mel front end read.  The kernels specialise on these (d_model 256, 4 heads x 64, FFN 1024 with k=9/1, 80 mels,
n_fft 1024 / hop 256); the Engine validates shapes at pack time."""
# Speaker embedding
speaker_embed_dim = 512
# Quantization for F0 and energy
f0_min = 71.0
f0_max = 797.9
energy_min = 0.1
energy_max = 525.43
# Audio and mel
sampling_rate = 22050
filter_length = 1024
hop_length = 256
win_length = 1024
n_bins = 256
max_wav_value = 32768.0
n_mel_channels = 80
mel_fmin = 0.0
mel_fmax = 8000.0
# STYLER
encoder_layer = 2
encoder_head = 4
encoder_hidden = 256
decoder_layer = 4
decoder_head = 4
decoder_hidden = 256
fft_conv1d_filter_size = 1024
fft_conv1d_kernel_size = (9, 1)
encoder_dropout = 0.2
decoder_dropout = 0.2
style_predictor_filter_size = 256
style_predictor_kernel_size = 3
style_predictor_dropout = 0.5
max_seq_len = 1000
dat_weight = 1
max_mel_len = 1024
va_neck_hidden_t = 4
va_neck_hidden_r = 64
va_neck_hidden_d = 80
va_neck_hidden_p = 64
va_neck_hidden_e = 64
va_enc_dim_r = 256
va_enc_dim_d = 256
va_enc_dim_p = 320
va_enc_dim_e = 320
va_dim_f0 = 257
va_dim_energy = 257
va_chs_grp = 16
# Optimizer-side constants kept for completeness of the surface
batch_size = 16
# Log-scaled duration
log_offset = 1.
