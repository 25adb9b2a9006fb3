"""The reference's three network declarations, written against layers.py exactly as the reference
writes them against Keras (same layer order -> same auto-names -> weights map 1:1).

  unet()        /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:853-915
                (identical at task3_lung_segmentation_unet.py:850-912, task1_crossval_*:919-981 / 957-1019)
  unetpp()      /root/reference/Scripts/task1_unet_plus_plus.py:860-949
  classifier()  /root/reference/Scripts/task2_covid19_classifcation.py:747-778
"""
from .layers import (BatchNormalization, Conv2D, Conv2DTranspose, Dense, Dropout, Flatten, Graph, Input,
                     MaxPooling2D, concatenate, reset_names)


def unet(new_dim=224, channels=1):
    reset_names()
    inputs = Input((new_dim, new_dim, channels))
    skips, x = [], inputs
    for ch in (32, 64, 128, 256):
        c = Conv2D(ch, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(x)
        c = Conv2D(ch, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(c)
        c = BatchNormalization()(c)
        skips.append(c)
        p = MaxPooling2D((2, 2))(c)
        x = Dropout(0.25)(p)
    c = Conv2D(512, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(x)
    c = Conv2D(512, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(c)
    for ch, skip in zip((256, 128, 64, 32), reversed(skips)):
        u = Conv2DTranspose(ch, (2, 2), strides=(2, 2), padding='same')(c)
        u = concatenate([u, skip], axis=3)
        u = BatchNormalization()(u)
        c = Conv2D(ch, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(u)
        c = Conv2D(ch, (3, 3), activation='relu', padding='same', kernel_initializer="he_normal")(c)
    outputs = Conv2D(1, (1, 1), activation='sigmoid')(c)
    return Graph(inputs=[inputs], outputs=[outputs])


def unetpp(new_dim=224, channels=1):
    reset_names()
    dropout_rate, activation = 0.4, "elu"

    def conv_block(t, ch):
        for _ in range(2):
            t = Conv2D(ch, (3, 3), activation=activation, kernel_initializer='he_normal', padding='same')(t)
            t = Dropout(dropout_rate)(t)
            t = BatchNormalization()(t)
        return t

    def backbone(t, ch):
        c = Conv2D(ch, (3, 3), activation='elu', kernel_initializer='he_normal', padding='same')(t)
        c = Dropout(0.2)(c)
        c = Conv2D(ch, (3, 3), activation='elu', kernel_initializer='he_normal', padding='same')(c)
        c = BatchNormalization()(c)
        return c, MaxPooling2D((2, 2))(c)

    inputs = Input((new_dim, new_dim, channels))
    c1, p1 = backbone(inputs, 32)
    c2, p2 = backbone(p1, 64)
    up1_2 = Conv2DTranspose(32, (2, 2), strides=(2, 2), padding='same')(c2)
    conv1_2 = conv_block(concatenate([up1_2, c1], axis=3), 32)
    c3, p3 = backbone(p2, 128)
    up2_2 = Conv2DTranspose(64, (2, 2), strides=(2, 2), padding='same')(c3)
    conv2_2 = conv_block(concatenate([up2_2, c2], axis=3), 64)
    up1_3 = Conv2DTranspose(32, (2, 2), strides=(2, 2), padding='same')(conv2_2)
    conv1_3 = conv_block(concatenate([up1_3, c1, conv1_2], axis=3), 32)
    c4, _p4 = backbone(p3, 256)          # p4 is declared but unused in the reference (UPP:912)
    up3_2 = Conv2DTranspose(128, (2, 2), strides=(2, 2), padding='same')(c4)
    conv3_2 = conv_block(concatenate([up3_2, c3], axis=3), 128)
    up2_3 = Conv2DTranspose(64, (2, 2), strides=(2, 2), padding='same')(conv3_2)
    conv2_3 = conv_block(concatenate([up2_3, c2, conv2_2], axis=3), 64)
    up1_4 = Conv2DTranspose(32, (2, 2), strides=(2, 2), padding='same')(conv2_3)
    conv1_4 = conv_block(concatenate([up1_4, c1, conv1_2, conv1_3], axis=3), 32)
    out = Conv2D(1, (1, 1), activation='sigmoid', kernel_initializer='he_normal', padding='same')(conv1_4)
    return Graph(inputs=[inputs], outputs=[out])


def classifier(new_dim=224, channels=1):
    reset_names()
    x = inputs = Input((new_dim, new_dim, channels))
    for ch in (16, 32, 64):
        x = Conv2D(ch, (3, 3), activation='relu', padding="same", kernel_initializer="he_normal")(x)
        x = BatchNormalization()(x)
        x = Conv2D(ch, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal")(x)
        x = BatchNormalization()(x)
        x = MaxPooling2D(pool_size=(2, 2))(x)
    x = Flatten()(x)
    x = Dense(32, activation='relu')(x)
    x = Dropout(0.4)(x)
    x = Dense(1, activation='sigmoid')(x)
    return Graph(inputs=[inputs], outputs=[x])


GRAPHS = {"unet": unet, "unetpp": unetpp, "classifier": classifier}
